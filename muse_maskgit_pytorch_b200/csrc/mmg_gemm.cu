// mmg_linear / mmg_conv2d / mmg_conv_transpose2d: launchers for the tcgen05 GEMM (bf16) and the fp32 CUDA-core GEMM.
#include "mmg_gemm_tc.cuh"
#include "mmg_tmap.cuh"
#include <mutex>

namespace mmg {

int linear_impl(const mmg_linear_args* a, const int* skip_if_zero, void* stream);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN>
static int launch_tc(const TcGemmParams& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return fail(MMG_ECUDA, "cudaFuncSetAttribute(tc_gemm<%d>): %s", BN, cudaGetErrorString(attr_err));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  MMG_CUDA(launch_pdl(tc_gemm_kernel<BN>, dim3(grid), dim3(TC_THREADS), TcCfg<BN>::SMEM_BYTES, st, p));
  MMG_LAUNCHED();
  return MMG_OK;
}

// in-place reduction epilogue (x += A W^T through cp.reduce.async.bulk): the residual GEMMs of the transformer blocks
template <int BN>
static int launch_tc_red(const TcGemmParams& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_gemm_kernel<BN, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES_RED); });
  if (attr_err != cudaSuccess) return fail(MMG_ECUDA, "cudaFuncSetAttribute(tc_gemm_red<%d>): %s", BN, cudaGetErrorString(attr_err));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  MMG_CUDA(launch_pdl(tc_gemm_kernel<BN, false, false, 2>, dim3(grid), dim3(TC_THREADS), TcCfg<BN>::SMEM_BYTES_RED, st, p));
  MMG_LAUNCHED();
  return MMG_OK;
}

// plain fp32 store through the per-warp tiles + TMA store (the logits GEMM)
template <int BN>
static int launch_tc_tstore(const TcGemmParams& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_gemm_kernel<BN, false, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES_RED); });
  if (attr_err != cudaSuccess) return fail(MMG_ECUDA, "cudaFuncSetAttribute(tc_gemm_tstore<%d>): %s", BN, cudaGetErrorString(attr_err));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  MMG_CUDA(launch_pdl(tc_gemm_kernel<BN, false, false, 3>, dim3(grid), dim3(TC_THREADS), TcCfg<BN>::SMEM_BYTES_RED, st, p));
  MMG_LAUNCHED();
  return MMG_OK;
}

// QKV (MODE 4) / GEGLU (MODE 5) epilogues through per-warp tiles + TMA stores
template <int BN, int MODE = 4>
static int launch_tc_qkvt(const TcGemmParams& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_gemm_kernel<BN, false, false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES_RED); });
  if (attr_err != cudaSuccess) return fail(MMG_ECUDA, "cudaFuncSetAttribute(tc_gemm_tiles<%d, %d>): %s", BN, MODE, cudaGetErrorString(attr_err));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  MMG_CUDA(launch_pdl(tc_gemm_kernel<BN, false, false, MODE>, dim3(grid), dim3(TC_THREADS), TcCfg<BN>::SMEM_BYTES_RED, st, p));
  MMG_LAUNCHED();
  return MMG_OK;
}

// LayerNorm-fused variant: clusters of two CTAs (column halves of the same rows), grid = 2 * min(#m-tiles, #SM / 2)
template <int BN>
static int launch_tc_lnf(const TcGemmParams& p, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_gemm_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return fail(MMG_ECUDA, "cudaFuncSetAttribute(tc_gemm_lnf<%d>): %s", BN, cudaGetErrorString(attr_err));
  const int pairs = p.num_m_tiles < num_sms() / 2 ? p.num_m_tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TcCfg<BN>::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MMG_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, true>, p));
  MMG_LAUNCHED();
  return MMG_OK;
}

// CTA-pair variant (cta_group::2): clusters of two CTAs, one 256 x BN tile per pair and step
template <int BN, int EPI_MODE = 0>
static int launch_tc_pair(const TcGemmParams& p, cudaStream_t st) {
  constexpr int SMEM = EPI_MODE ? TcCfg<BN>::PAIR_SMEM_BYTES_RED : TcCfg<BN>::PAIR_SMEM_BYTES;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  static int max_clusters = 0;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(tc_gemm_kernel<BN, false, true, EPI_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (attr_err != cudaSuccess) return;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (num_sms() / 2)); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    attr_err = cudaOccupancyMaxActiveClusters(&max_clusters, tc_gemm_kernel<BN, false, true, EPI_MODE>, &cfg);
  });
  if (attr_err != cudaSuccess || max_clusters < 1) return fail(MMG_ECUDA, "tc_gemm pair<%d> setup: %s (clusters %d)", BN, cudaGetErrorString(attr_err), max_clusters);
  const int tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MMG_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, false, true, EPI_MODE>, p));
  MMG_LAUNCHED();
  return MMG_OK;
}

// CTA pairs (cta_group::2) cut the L2->SM operand traffic by a third and deepen the ring to six stages: 8192^3 788 -> 732 us (88 % of the
// measured cuBLAS rate), the VAE 3x3 convolutions (K = 9 Cin) -10 %, FF2 (K = 1408) 57 -> 51 us.  Round 2: the epilogue warps used to hand
// the TMEM stage back to the leader CTA with mbarrier.arrive.release.cluster, which compiles to MEMBAR.ALL.GPU + ERRBAR and made every
// warp wait, once per tile, for all of its global stores; with the plain remote arrive (mbar_arrive_remote) pairs also win for the K = 512
// GEMMs with register epilogues (QKV 54 -> 48 us, GEGLU FF1 101 -> 89 us, fp32 logits 741 -> 640 us) and are the default for every dense
// product with 256-column tiles; the short residual GEMM (N = 512, two column tiles) stays on single CTAs.  MMG_GEMM_PAIR = 0 / 1 forces the choice.
static bool use_pair(const TcGemmParams& p, int bn, int epi_mode) {
  static const int forced = [] { const char* e = getenv("MMG_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  if (bn != 256 || forced == 0 || p.num_m_tiles < 2) return false;
  if (forced == 1) return true;
  if (((p.num_m_tiles + 1) / 2) * p.num_n_tiles < 64) return false;
  if (epi_mode == 3) return true;
  if (epi_mode == 2) return p.num_kb >= 16;
  if (p.mode == 0) return true;                                  // dense products: QKV, GEGLU, plain stores
  return p.num_kb >= 64 && p.epi.kind != MMG_EPI_CONVT && p.epi.kind != MMG_EPI_CONVT_RGB;
}

static int pick_bn(int64_t M_tiles, int64_t N, int epilogue) {
  if (epilogue == MMG_EPI_CONVT_RGB) return (int)N;             // whole row in one tile
  if (N % 256 == 0 && M_tiles * (N / 256) >= 2 * 148) return 256;
  if (N % 128 == 0 && M_tiles * (N / 128) >= 148) return 128;
  if (N % 128 == 0 && N >= 1024) return 128;
  return (N % 64 == 0 && N % 128 != 0) ? 64 : ((M_tiles * ((N + 127) / 128) >= 96) ? 128 : 64);
}

static int dispatch_tc(TcGemmParams& p, int bn, const void* w, int64_t N, int64_t K, int64_t ldw, cudaStream_t st) {
  p.num_n_tiles = (int)((N + bn - 1) / bn);
  {
    static const int nfast_forced = [] { const char* e = getenv("MMG_GEMM_NFAST"); return e ? atoi(e) : -1; }();
    p.n_fast = nfast_forced >= 0 ? nfast_forced : ((int64_t)N * K * 2 <= (8 << 20) && p.num_n_tiles > 1 && p.num_n_tiles <= 16) ? 1 : 0;
  }
  const mmg_epilogue_args& e = p.epi.p;
  const bool in_place = (p.epi.kind == MMG_EPI_LNFOLD_RESIDUAL || (p.epi.kind == MMG_EPI_RESIDUAL && e.act == 0)) && e.out_dtype == MMG_F32 &&
                        e.out == e.resid && e.ldo == e.ldr && (e.ldo % 4) == 0 && aligned16(e.out) && !e.ln_out && p.mode == 0;
  const bool plain_f32 = p.epi.kind == MMG_EPI_STORE && e.out_dtype == MMG_F32 && p.mode == 0 && !e.bias && e.act == 0 && (e.ldo % 4) == 0 && aligned16(e.out);
  static const int red_forced = [] { const char* ev = getenv("MMG_GEMM_RED"); return ev ? atoi(ev) : -1; }();
  static const int tstore_forced = [] { const char* ev = getenv("MMG_GEMM_TSTORE"); return ev ? atoi(ev) : -1; }();
  static const int qkvt_forced = [] { const char* ev = getenv("MMG_GEMM_QKVT"); return ev ? atoi(ev) : -1; }();
  const bool qkv_tiles = p.epi.kind == MMG_EPI_QKV && e.out_dtype == MMG_BF16 && p.mode == 0 && e.tokens % 32 == 0 && p.M % 128 == 0 && (bn == 128 || bn == 256) &&
                         (!e.nq_heads || aligned16(e.q_out)) && (!e.nk_heads || (aligned16(e.k_out) && aligned16(e.v_out))) && e.nk_heads == e.nv_heads;
  static const int geglut_forced = [] { const char* ev = getenv("MMG_GEMM_GEGLUT"); return ev ? atoi(ev) : -1; }();
  const bool geglu_tiles = p.epi.kind == MMG_EPI_GEGLU && e.out_dtype == MMG_BF16 && p.mode == 0 && bn == 256 && (e.ldo % 8) == 0 && aligned16(e.out);
  const int epi_mode = (in_place && red_forced != 0) ? 2 : (plain_f32 && tstore_forced != 0 && (bn == 256 || bn == 128)) ? 3 : (qkv_tiles && qkvt_forced != 0) ? 4 :
                       (geglu_tiles && geglut_forced != 0) ? 5 : 0;
  const bool pair = use_pair(p, bn, epi_mode >= 4 ? 0 : epi_mode);
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}; uint64_t str[1] = {(uint64_t)ldw * 2}; uint32_t box[2] = {TC_BK, (uint32_t)(pair ? bn / 2 : bn)};
  int rc = make_tmap_bf16(&p.tma_b, w, 2, dims, str, box); if (rc) return rc;
  if (epi_mode == 4) {
    const uint64_t seqs = (uint64_t)(p.M / e.tokens) * (uint64_t)e.heads;
    uint64_t str[1] = {128}; uint32_t box[2] = {64, 32};
    if (e.nq_heads) { uint64_t d[2] = {64, seqs * (uint64_t)e.q_rows}; rc = make_tmap_bf16(&p.tma_qkv[0], e.q_out, 2, d, str, box); if (rc) return rc; }
    if (e.nk_heads) {
      uint64_t d[2] = {64, seqs * (uint64_t)e.kv_rows};
      rc = make_tmap_bf16(&p.tma_qkv[1], e.k_out, 2, d, str, box); if (rc) return rc;
      rc = make_tmap_bf16(&p.tma_qkv[2], e.v_out, 2, d, str, box); if (rc) return rc;
    }
    if (!pair) return bn == 256 ? launch_tc_qkvt<256>(p, st) : launch_tc_qkvt<128>(p, st);
    return launch_tc_pair<256, 4>(p, st);
  }
  if (epi_mode == 5) {
    uint64_t od[2] = {(uint64_t)(p.N / 2), (uint64_t)p.M}; uint64_t os[1] = {(uint64_t)e.ldo * 2}; uint32_t ob[2] = {64, 32};
    rc = make_tmap_bf16(&p.tma_out, e.out, 2, od, os, ob); if (rc) return rc;
    return pair ? launch_tc_pair<256, 5>(p, st) : launch_tc_qkvt<256, 5>(p, st);
  }
  if (epi_mode) {
    uint64_t od[2] = {(uint64_t)p.N, (uint64_t)p.M}; uint64_t os[1] = {(uint64_t)e.ldo * 4}; uint32_t ob[2] = {32, 32};
    rc = make_tmap_f32(&p.tma_out, e.out, 2, od, os, ob); if (rc) return rc;
  }
  if (pair) return epi_mode == 2 ? launch_tc_pair<256, 2>(p, st) : epi_mode == 3 ? launch_tc_pair<256, 3>(p, st) : launch_tc_pair<256>(p, st);
  if (epi_mode == 3) return bn == 256 ? launch_tc_tstore<256>(p, st) : launch_tc_tstore<128>(p, st);   // 128: the sample GEMM of the fused tail at small batch
  if (epi_mode == 2) {
    switch (bn) { case 64: return launch_tc_red<64>(p, st); case 128: return launch_tc_red<128>(p, st); case 256: return launch_tc_red<256>(p, st); }
  }
  switch (bn) {
    case 64: return launch_tc<64>(p, st);
    case 128: return launch_tc<128>(p, st);
    case 256: return launch_tc<256>(p, st);
  }
  return fail(MMG_EINVAL, "unsupported tile width %d", bn);
}

// ------------------------------------------------------------------------------------------------------------------
// fp32-accumulate CUDA-core GEMM (parity precision, and shapes the TMA path does not take).  64x64x16 tiles,
// 256 threads, 4x4 micro-tile per thread; A is addressed through a loader (dense or implicit conv).
// ------------------------------------------------------------------------------------------------------------------
struct ConvGeom {
  int B, H, W, Cin, Ho, Wo, stride, ntaps;
  int8_t tdy[25], tdx[25];
};

template <typename T, bool CONV>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const T* __restrict__ A, const T* __restrict__ Wt, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw,
                 ConvGeom g, Epilogue epi) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  __shared__ float Cs[64][64 + 1];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  // loader mapping: each thread loads 4 elements of A and of W per k-step: row = tid/4 (0..63), k = (tid%4)*4 .. +3
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int64_t am = m0 + lr;
  int ab = 0, ay = 0, ax = 0;
  if (CONV && am < M) { ab = (int)(am / ((int64_t)g.Ho * g.Wo)); int rem = (int)(am - (int64_t)ab * g.Ho * g.Wo); ay = rem / g.Wo; ax = rem - ay * g.Wo; }
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + lk + j;
      float av = 0.f, bv = 0.f;
      if (am < M && k < K) {
        if (!CONV) av = to_f(A[am * lda + k]);
        else {
          const int tap = (int)(k / g.Cin), c = (int)(k - (int64_t)tap * g.Cin);
          const int iy = ay * g.stride + g.tdy[tap], ix = ax * g.stride + g.tdx[tap];
          if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) av = to_f(A[(((int64_t)ab * g.H + iy) * g.W + ix) * g.Cin + c]);
        }
      }
      if (n0 + lr < N && k < K) bv = to_f(Wt[(int64_t)(n0 + lr) * ldw + k]);
      As[lk + j][lr] = av; Bs[lk + j][lr] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  if (tid < 64) {
    const int64_t row = m0 + tid;
    if (row < M) {
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = Cs[tid][i];
      int nvalid = (int)(N - n0); if (nvalid > 64) nvalid = 64;
      // CONVT_RGB needs the whole row in one pass and is not offered on this path (validated by the launcher)
      epi.begin_row(row);
      epi.template apply<false>(row, n0, v, nvalid);
      epi.end_row(row);
    }
  }
}

template <bool CONV>
static int launch_simt(int dtype, const void* a, const void* w, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw,
                       const ConvGeom& g, const Epilogue& epi, cudaStream_t st) {
  MMG_CHECK_ARG(epi.kind != MMG_EPI_CONVT_RGB, "CONVT_RGB epilogue requires the bf16 tensor-core path");
  MMG_CHECK_ARG(M > 0 && N > 0 && K > 0, "empty GEMM");
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
  MMG_CHECK_ARG(grid.y < 65536, "N too large for the fp32 path");
  if (dtype == MMG_F32) simt_gemm_kernel<float, CONV><<<grid, 256, 0, st>>>((const float*)a, (const float*)w, M, N, K, lda, ldw, g, epi);
  else simt_gemm_kernel<bf16, CONV><<<grid, 256, 0, st>>>((const bf16*)a, (const bf16*)w, M, N, K, lda, ldw, g, epi);
  simt_launch_counter()++;
  MMG_LAUNCHED();
  return MMG_OK;
}


// conv tap tables.  kind 0: 1x1; 1: 3x3 s1 p1; 2: 4x4 s2 p1; 3: 5x5 s1 p2.  Tap index t = r*kw + s (weights packed tap-major).
static void conv_geom(int kind, int B, int H, int W, int Cin, ConvGeom* g) {
  const int k = kind == 0 ? 1 : kind == 1 ? 3 : kind == 2 ? 4 : 5;
  const int pad = kind == 0 ? 0 : kind == 1 ? 1 : kind == 2 ? 1 : 2;
  g->B = B; g->H = H; g->W = W; g->Cin = Cin; g->stride = (kind == 2) ? 2 : 1;
  g->Ho = H / g->stride; g->Wo = W / g->stride; g->ntaps = k * k;
  for (int r = 0; r < k; ++r) for (int s = 0; s < k; ++s) { g->tdy[r * k + s] = (int8_t)(r - pad); g->tdx[r * k + s] = (int8_t)(s - pad); }
}

// conv-transpose (k4 s2 p1) parity class (py,px): 2x2 taps, tap t = a*2+b, input offset dy = DY[py][a], kernel row r = R[py][a].
static const int CT_D[2][2] = {{0, -1}, {1, 0}};

static int tile_geometry(int Ho, int Wo, TcGemmParams* p) {
  int TW = Wo < 128 ? Wo : 128;
  if (128 % TW != 0 || Wo % TW != 0) return -1;
  int TH = 128 / TW; if (TH > Ho) TH = Ho;
  if (Ho % TH != 0 || 128 % (TW * TH) != 0) return -1;
  p->TW = TW; p->TH = TH; p->TB = 128 / (TW * TH);
  p->tiles_x = Wo / TW; p->tiles_y = Ho / TH;
  return 0;
}

}  // namespace mmg

using namespace mmg;

extern "C" int mmg_linear(const mmg_linear_args* a, void* stream) { return mmg::linear_impl(a, nullptr, stream); }

// skip_if_zero: optional device word; the tensor-core kernel returns at once when it reads 0 there (the fallback leg of mmg_logits_fused)
int mmg::linear_impl(const mmg_linear_args* a, const int* skip_if_zero, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMG_CHECK_ARG(a && a->a && a->w, "mmg_linear: NULL operand");
  MMG_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "mmg_linear: bad shape M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K);
  int rc = validate_epilogue(a->epilogue, a->epi, a->N); if (rc) return rc;
  Epilogue epi; epi.p = a->epi; epi.kind = a->epilogue; epi.fast = 0; epi.M = a->M; epi.N = a->N;
  const bool tc_ok = a->dtype == MMG_BF16 && (a->K % 64 == 0) && (a->N % 64 == 0) && (a->lda % 8 == 0) && (a->ldw % 8 == 0) &&
                     aligned16(a->a) && aligned16(a->w) && (a->epi.ldo % 8 == 0 || a->epilogue == MMG_EPI_QKV || a->epilogue == MMG_EPI_CONVT_RGB || a->epilogue == MMG_EPI_LFQ_IDS || a->epilogue == MMG_EPI_ARGMIN);
  if (!tc_ok) {
    MMG_CHECK_ARG(!skip_if_zero, "skippable GEMM requires the bf16 tensor-core path");
    if (a->dtype == MMG_BF16) note_simt_fallback("mmg_linear", a->M, a->N, a->K);
    MMG_CHECK_ARG(!a->epi.ln_out, "fused LayerNorm output requires the bf16 tensor-core path (K %% 64, N %% 64, aligned operands)");
    ConvGeom g{};
    return launch_simt<false>(a->dtype, a->a, a->w, a->M, a->N, a->K, a->lda, a->ldw, g, epi, st);
  }
  TcGemmParams p{};
  p.skip_if_zero = skip_if_zero;
  p.M = a->M; p.N = a->N; p.num_kb = (int)(a->K / TC_BK); p.mode = 0;
  p.num_m_tiles = (int)((a->M + TC_BM - 1) / TC_BM);
  p.epi = epi; p.epi.fast = 1;
  uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M}; uint64_t str[1] = {(uint64_t)a->lda * 2}; uint32_t box[2] = {TC_BK, TC_BM};
  rc = make_tmap_bf16(&p.tma_a[0], a->a, 2, dims, str, box); if (rc) return rc;
  if (a->epi.ln_out) {
    const int bn = (int)(a->N / 2);
    MMG_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "fused LayerNorm output needs N in {128, 256, 512} (two column tiles), got %lld", (long long)a->N);
    MMG_CHECK_ARG((a->epilogue == MMG_EPI_RESIDUAL && a->epi.out_dtype == MMG_F32 && a->epi.act == 0 && a->epi.ldr % 8 == 0) || a->epilogue == MMG_EPI_LNFOLD_RESIDUAL,
                  "fused LayerNorm output is offered for the fp32 RESIDUAL / LNFOLD_RESIDUAL epilogues");
    MMG_CHECK_ARG(a->epi.ln_gamma && a->epi.ld_ln % 8 == 0, "fused LayerNorm output: gamma / ld_ln");
    uint64_t dimsb[2] = {(uint64_t)a->K, (uint64_t)a->N}; uint64_t strb[1] = {(uint64_t)a->ldw * 2}; uint32_t boxb[2] = {TC_BK, (uint32_t)bn};
    rc = make_tmap_bf16(&p.tma_b, a->w, 2, dimsb, strb, boxb); if (rc) return rc;
    p.num_n_tiles = 2;
    switch (bn) { case 64: return launch_tc_lnf<64>(p, st); case 128: return launch_tc_lnf<128>(p, st); default: return launch_tc_lnf<256>(p, st); }
  }
  const int bn = pick_bn(p.num_m_tiles, a->N, a->epilogue);
  MMG_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "mmg_linear: unsupported N=%lld for this epilogue", (long long)a->N);
  return dispatch_tc(p, bn, a->w, a->N, a->K, a->ldw, st);
}

extern "C" int mmg_conv2d(const mmg_conv2d_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMG_CHECK_ARG(a && a->x && a->w, "mmg_conv2d: NULL operand");
  MMG_CHECK_ARG(a->kind >= 0 && a->kind <= 3, "mmg_conv2d: kind %d", a->kind);
  MMG_CHECK_ARG(a->kind != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "mmg_conv2d: stride-2 conv needs even H, W");
  ConvGeom g; conv_geom(a->kind, a->B, a->H, a->W, a->Cin, &g);
  const int64_t M = (int64_t)a->B * g.Ho * g.Wo, N = a->Cout, K = (int64_t)g.ntaps * a->Cin;
  int rc = validate_epilogue(a->epilogue, a->epi, N); if (rc) return rc;
  Epilogue epi; epi.p = a->epi; epi.kind = a->epilogue; epi.fast = 0; epi.M = M; epi.N = N;
  TcGemmParams p{};
  const bool tc_ok = a->dtype == MMG_BF16 && a->kind != 3 && (a->Cin % 64 == 0) && (N % 64 == 0) && aligned16(a->x) && aligned16(a->w) &&
                     (a->epi.ldo % 8 == 0) && tile_geometry(g.Ho, g.Wo, &p) == 0;
  if (!tc_ok) {
    if (a->dtype == MMG_BF16) note_simt_fallback("mmg_conv2d", M, N, K);
    return launch_simt<true>(a->dtype, a->x, a->w, M, N, K, 0, K, g, epi, st);
  }
  p.M = M; p.N = N; p.mode = 1; p.cchunks = a->Cin / 64; p.ntaps = g.ntaps; p.num_kb = p.ntaps * p.cchunks;
  p.Ho = g.Ho; p.Wo = g.Wo; p.B = a->B;
  p.num_m_tiles = p.tiles_x * p.tiles_y * ((a->B + p.TB - 1) / p.TB);
  p.epi = epi; p.epi.fast = 1;
  const uint64_t C = a->Cin, W = a->W, H = a->H;
  uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TB};
  if (a->kind != 2) {
    uint64_t dims[4] = {C, W, H, (uint64_t)a->B}; uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    rc = make_tmap_bf16(&p.tma_a[0], a->x, 4, dims, str, box); if (rc) return rc;
    for (int t = 0; t < g.ntaps; ++t) { p.tap_map[t] = 0; p.tap_dy[t] = g.tdy[t]; p.tap_dx[t] = g.tdx[t]; }
  } else {
    // stride 2: input row 2y + (r-1) = 2(y+dy) + py.  One tensor map per input parity (py,px) over the (H/2, W/2) grid.
    for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px) {
      const uint8_t* base = reinterpret_cast<const uint8_t*>(a->x) + ((uint64_t)py * W + px) * C * 2;
      uint64_t dims[4] = {C, W / 2, H / 2, (uint64_t)a->B}; uint64_t str[3] = {2 * C * 2, 2 * W * C * 2, H * W * C * 2};
      rc = make_tmap_bf16(&p.tma_a[py * 2 + px], base, 4, dims, str, box); if (rc) return rc;
    }
    for (int r = 0; r < 4; ++r) for (int s = 0; s < 4; ++s) {
      const int oy = r - 1, ox = s - 1;
      const int py = ((oy % 2) + 2) % 2, px = ((ox % 2) + 2) % 2;
      p.tap_map[r * 4 + s] = (int8_t)(py * 2 + px); p.tap_dy[r * 4 + s] = (int8_t)((oy - py) / 2); p.tap_dx[r * 4 + s] = (int8_t)((ox - px) / 2);
    }
  }
  const int bn = pick_bn(p.num_m_tiles, N, a->epilogue);
  return dispatch_tc(p, bn, a->w, N, K, K, st);
}

extern "C" int mmg_conv_transpose2d(const mmg_conv_transpose2d_args* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MMG_CHECK_ARG(a && a->x && a->w, "mmg_conv_transpose2d: NULL operand");
  MMG_CHECK_ARG(a->epilogue == MMG_EPI_CONVT || a->epilogue == MMG_EPI_CONVT_RGB, "mmg_conv_transpose2d: epilogue must be CONVT/CONVT_RGB");
  const int64_t M = (int64_t)a->B * a->H * a->W, N = a->Cout, K = 4 * (int64_t)a->Cin;
  MMG_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "mmg_conv_transpose2d: bad geometry");
  int rc;
  { mmg_epilogue_args e = a->epi; e.H = a->H; e.W = a->W; rc = validate_epilogue(a->epilogue, e, N); if (rc) return rc; }
  TcGemmParams p{};
  const bool tc_ok = a->dtype == MMG_BF16 && (a->Cin % 64 == 0) && (N % 64 == 0) && aligned16(a->x) && aligned16(a->w) &&
                     (a->epilogue == MMG_EPI_CONVT_RGB ? (N == 64 || N == 128 || N == 256) : (a->epi.ldo % 8 == 0)) &&
                     tile_geometry(a->H, a->W, &p) == 0;
  const size_t esz = a->dtype == MMG_BF16 ? 2 : 4;
  if (!tc_ok && a->dtype == MMG_BF16) note_simt_fallback("mmg_conv_transpose2d", M, N, K);
  for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px) {
    Epilogue epi; epi.p = a->epi; epi.kind = a->epilogue; epi.fast = 0; epi.M = M; epi.N = N;
    epi.p.H = a->H; epi.p.W = a->W; epi.p.py = py; epi.p.px = px;
    const uint8_t* wp = reinterpret_cast<const uint8_t*>(a->w) + (size_t)(py * 2 + px) * N * K * esz;
    if (!tc_ok) {
      ConvGeom g{}; g.B = a->B; g.H = a->H; g.W = a->W; g.Cin = a->Cin; g.Ho = a->H; g.Wo = a->W; g.stride = 1; g.ntaps = 4;
      for (int t = 0; t < 4; ++t) { g.tdy[t] = (int8_t)CT_D[py][t >> 1]; g.tdx[t] = (int8_t)CT_D[px][t & 1]; }
      rc = launch_simt<true>(a->dtype, a->x, wp, M, N, K, 0, K, g, epi, st); if (rc) return rc;
      continue;
    }
    TcGemmParams q = p;
    q.M = M; q.N = N; q.mode = 1; q.cchunks = a->Cin / 64; q.ntaps = 4; q.num_kb = 4 * q.cchunks;
    q.Ho = a->H; q.Wo = a->W; q.B = a->B;
    q.num_m_tiles = q.tiles_x * q.tiles_y * ((a->B + q.TB - 1) / q.TB);
    q.epi = epi; q.epi.fast = 1;
    const uint64_t C = a->Cin, W = a->W, H = a->H;
    uint64_t dims[4] = {C, W, H, (uint64_t)a->B}; uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)q.TW, (uint32_t)q.TH, (uint32_t)q.TB};
    rc = make_tmap_bf16(&q.tma_a[0], a->x, 4, dims, str, box); if (rc) return rc;
    for (int t = 0; t < 4; ++t) { q.tap_map[t] = 0; q.tap_dy[t] = (int8_t)CT_D[py][t >> 1]; q.tap_dx[t] = (int8_t)CT_D[px][t & 1]; }
    const int bn = pick_bn(q.num_m_tiles, N, a->epilogue);
    rc = dispatch_tc(q, bn, wp, N, K, K, st); if (rc) return rc;
  }
  return MMG_OK;
}

#ifdef MMG_GEMM_TRACE
extern "C" int mmg_trace_clear(void) {
  void* p = nullptr;
  if (cudaGetSymbolAddress(&p, mmg::g_gemm_trace) != cudaSuccess) return -1;
  return (int)cudaMemset(p, 0, sizeof(mmg::g_gemm_trace));
}
extern "C" int mmg_trace_read(void* dst, size_t bytes) {
  return (int)cudaMemcpyFromSymbol(dst, mmg::g_gemm_trace, bytes < sizeof(mmg::g_gemm_trace) ? bytes : sizeof(mmg::g_gemm_trace));
}
#endif
