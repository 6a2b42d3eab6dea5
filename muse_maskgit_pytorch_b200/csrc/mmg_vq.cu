// VQ codebook lookups (HBM-bound scans).  replaces vqgan_vae.py:424 (quantizer forward) and 429-435 (ids -> codes).
#include "mmg_common.cuh"
#include <float.h>

namespace mmg {

// LFQ encode: the nearest code of the implicit {+-1}^bits codebook is the sign pattern of the projected token.
// One warp per token: 128-bit coalesced loads of the token's D channels, `bits` running dot products against the
// projection rows held in shared memory, warp-shuffle reduction, then bit-pack (MSB first; x == 0 -> bit 0).
template <typename T, int BITS_MAX>
__global__ void __launch_bounds__(256)
lfq_encode_kernel(const T* __restrict__ x, const float* __restrict__ w_in, const float* __restrict__ b_in, int64_t* __restrict__ ids,
                  int64_t tokens, int D, int bits) {
  extern __shared__ float ws[];                    // [bits][D]
  if (w_in) for (int i = threadIdx.x; i < bits * D; i += blockDim.x) ws[i] = w_in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < tokens; t += (int64_t)gridDim.x * 8) {
    const T* xr = x + t * D;
    float acc[BITS_MAX];
#pragma unroll
    for (int i = 0; i < BITS_MAX; ++i) acc[i] = 0.f;
    if (w_in) {
      for (int c = lane; c < D; c += 32) {
        const float xv = to_f(xr[c]);
#pragma unroll
        for (int i = 0; i < BITS_MAX; ++i) if (i < bits) acc[i] = fmaf(xv, ws[i * D + c], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < BITS_MAX; ++i) { acc[i] = warp_sum(acc[i]); if (i < bits && b_in) acc[i] += b_in[i]; }
    } else {
#pragma unroll
      for (int i = 0; i < BITS_MAX; ++i) if (i < bits) acc[i] = to_f(xr[i]);
    }
    if (lane == 0) {
      int64_t id = 0;
#pragma unroll
      for (int i = 0; i < BITS_MAX; ++i) if (i < bits && acc[i] > 0.f) id |= (int64_t)1 << (bits - 1 - i);
      ids[t] = id;
    }
  }
}

// Explicit codebook: ids[t] = argmin_k ||x_t - e_k||^2 = argmin_k (||e_k||^2 - 2 x_t.e_k)  (+||x_t||^2, constant in k).
// CTA = 8 warps x 4 tokens per warp kept in shared memory; the codebook streams through once per CTA in coalesced
// 128-bit loads (code k is read by one warp, all 32 lanes across D); warp-shuffle sum, running (min, argmin) per
// token in registers (first index wins ties), cross-warp argmin at the end.
constexpr int VQ_TOK = 16;   // tokens per CTA
__global__ void __launch_bounds__(256)
vq_l2_argmin_kernel(const float* __restrict__ x, const float* __restrict__ cb, int64_t* __restrict__ ids, int64_t tokens, int K, int D) {
  extern __shared__ float xs[];                    // [VQ_TOK][D]
  __shared__ float bd[8][VQ_TOK]; __shared__ int bi[8][VQ_TOK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t0 = (int64_t)blockIdx.x * VQ_TOK;
  for (int i = threadIdx.x; i < VQ_TOK * D; i += blockDim.x) { const int64_t t = t0 + i / D; xs[i] = t < tokens ? x[t * D + (i % D)] : 0.f; }
  __syncthreads();
  float best[VQ_TOK]; int besti[VQ_TOK];
#pragma unroll
  for (int j = 0; j < VQ_TOK; ++j) { best[j] = FLT_MAX; besti[j] = 0x7fffffff; }
  for (int k = warp; k < K; k += 8) {
    const float* e = cb + (int64_t)k * D;
    float dot[VQ_TOK]; float e2 = 0.f;
#pragma unroll
    for (int j = 0; j < VQ_TOK; ++j) dot[j] = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float ev = __ldg(e + c);
      e2 = fmaf(ev, ev, e2);
#pragma unroll
      for (int j = 0; j < VQ_TOK; ++j) dot[j] = fmaf(ev, xs[j * D + c], dot[j]);
    }
    e2 = warp_sum(e2);
#pragma unroll
    for (int j = 0; j < VQ_TOK; ++j) {
      const float d = e2 - 2.f * warp_sum(dot[j]);
      if (d < best[j] || (d == best[j] && k < besti[j])) { best[j] = d; besti[j] = k; }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < VQ_TOK; ++j) { bd[warp][j] = best[j]; bi[warp][j] = besti[j]; }
  }
  __syncthreads();
  if (threadIdx.x < VQ_TOK) {
    const int j = threadIdx.x; float d = bd[0][j]; int i = bi[0][j];
    for (int w = 1; w < 8; ++w) if (bd[w][j] < d || (bd[w][j] == d && bi[w][j] < i)) { d = bd[w][j]; i = bi[w][j]; }
    if (t0 + j < tokens) ids[t0 + j] = i;
  }
}

// ids -> +-1 codes -> project_out (LFQ.indices_to_codes): out[t, c] = b_out[c] + sum_i (+-1)_i w_out[c, i]
template <typename T>
__global__ void __launch_bounds__(256)
vq_decode_codes_kernel(const int64_t* __restrict__ ids, const float* __restrict__ w_out, const float* __restrict__ b_out, T* __restrict__ out,
                       int64_t tokens, int D, int bits) {
  const int64_t t = blockIdx.x;
  const int64_t id = ids[t];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc;
    if (w_out) {
      acc = b_out ? b_out[c] : 0.f;
      for (int i = 0; i < bits; ++i) { const float w = __ldg(w_out + (int64_t)c * bits + i); acc += ((id >> (bits - 1 - i)) & 1) ? w : -w; }
    } else {
      acc = ((id >> (bits - 1 - c)) & 1) ? 1.f : -1.f;
    }
    out[t * D + c] = from_f<T>(acc);
  }
}

}  // namespace mmg

using namespace mmg;

extern "C" int mmg_vq_lfq_encode(const mmg_vq_lfq_encode_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMG_CHECK_ARG(a && a->x && a->ids, "mmg_vq_lfq_encode: NULL pointer");
  MMG_CHECK_ARG(a->bits >= 1 && a->bits <= 24, "mmg_vq_lfq_encode: bits=%d not in [1,24]", a->bits);
  MMG_CHECK_ARG(a->w_in || a->D == a->bits, "mmg_vq_lfq_encode: identity projection needs D == bits");
  if (a->T == 0) return MMG_OK;
  const size_t smem = a->w_in ? (size_t)a->bits * a->D * 4 : 0;
  MMG_CHECK_ARG(smem <= 200 * 1024, "mmg_vq_lfq_encode: projection does not fit in shared memory");
  int64_t grid = (a->T + 7) / 8; if (grid > (int64_t)num_sms() * 4) grid = (int64_t)num_sms() * 4;
  if (a->dtype == MMG_BF16) {
    MMG_CUDA(cudaFuncSetAttribute(lfq_encode_kernel<bf16, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lfq_encode_kernel<bf16, 24><<<(unsigned)grid, 256, smem, st>>>((const bf16*)a->x, a->w_in, a->b_in, a->ids, a->T, a->D, a->bits);
  } else {
    MMG_CUDA(cudaFuncSetAttribute(lfq_encode_kernel<float, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lfq_encode_kernel<float, 24><<<(unsigned)grid, 256, smem, st>>>((const float*)a->x, a->w_in, a->b_in, a->ids, a->T, a->D, a->bits);
  }
  MMG_LAUNCHED();
  return MMG_OK;
}

extern "C" int mmg_vq_l2_argmin(const mmg_vq_l2_argmin_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMG_CHECK_ARG(a && a->x && a->codebook && a->ids, "mmg_vq_l2_argmin: NULL pointer");
  MMG_CHECK_ARG(a->K >= 1 && a->D >= 1, "mmg_vq_l2_argmin: K, D");
  if (a->T == 0) return MMG_OK;
  const size_t smem = (size_t)VQ_TOK * a->D * 4;
  MMG_CHECK_ARG(smem <= 200 * 1024, "mmg_vq_l2_argmin: D too large");
  MMG_CUDA(cudaFuncSetAttribute(vq_l2_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vq_l2_argmin_kernel<<<(unsigned)((a->T + VQ_TOK - 1) / VQ_TOK), 256, smem, st>>>(a->x, a->codebook, a->ids, a->T, a->K, a->D);
  MMG_LAUNCHED();
  return MMG_OK;
}

extern "C" int mmg_vq_decode_codes(const mmg_vq_decode_codes_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMG_CHECK_ARG(a && a->ids && a->out, "mmg_vq_decode_codes: NULL pointer");
  MMG_CHECK_ARG(a->bits >= 1 && a->bits <= 62 && (a->w_out || a->D == a->bits), "mmg_vq_decode_codes: bits / identity projection");
  if (a->T == 0) return MMG_OK;
  if (a->dtype == MMG_BF16) vq_decode_codes_kernel<bf16><<<(unsigned)a->T, 256, 0, st>>>(a->ids, a->w_out, a->b_out, (bf16*)a->out, a->T, a->D, a->bits);
  else vq_decode_codes_kernel<float><<<(unsigned)a->T, 256, 0, st>>>(a->ids, a->w_out, a->b_out, (float*)a->out, a->T, a->D, a->bits);
  MMG_LAUNCHED();
  return MMG_OK;
}
