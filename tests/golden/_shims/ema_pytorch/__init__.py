"""Empty stand-in (trainers.py imports EMA; training is out of scope)."""
class EMA: pass
