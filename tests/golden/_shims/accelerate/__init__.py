"""Empty stand-in so `import muse_maskgit_pytorch` (which pulls trainers.py) succeeds offline."""
class Accelerator: pass
class DistributedType: pass
class DistributedDataParallelKwargs: pass
