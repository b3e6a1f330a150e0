// fp32 SIMT GEMM (see gemm_f32.cuh). Register-blocked, double-buffered shared-memory tiles.
// Three tile shapes cover the path: 128x128 (batched clips), 64x64 (mid), 32x32 (batch-1, M=90).
#include "gemm_f32.cuh"

namespace mocha {

namespace {

constexpr int BK = 16;

__device__ __forceinline__ int reflect_idx(int t, int T) {
  if (t < 0) t = -t;
  if (t >= T) t = 2 * (T - 1) - t;
  return t;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const GemmParams p) {
  pdl_trigger();
  pdl_wait();
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TXN = BN / TN;                 // threads along N
  constexpr int AV = (BM * BK) / (4 * NT);     // float4 loads of A per thread
  constexpr int BV = (BN * BK) / (4 * NT);     // float4 loads of W per thread
  static_assert(AV >= 1 && BV >= 1, "tile too small for the thread count");
  static_assert(TM == 4 || TM == 8, "TM");
  static_assert(TN == 4 || TN == 8, "TN");

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const int z1 = blockIdx.z / p.nz2, z2 = blockIdx.z % p.nz2;

  const float* __restrict__ A = p.A + z1 * p.sA1 + z2 * p.sA2;
  const float* __restrict__ W = p.W + z1 * p.sW1 + z2 * p.sW2;
  float* __restrict__ C = p.C + z1 * p.sC1 + z2 * p.sC2;
  const float* __restrict__ R = p.res ? p.res + z1 * p.sR1 + z2 * p.sR2 : nullptr;

  const bool a_vec = ((p.lda & 3) == 0) && ((p.K & 3) == 0) && ((((uintptr_t)A) & 15) == 0) &&
                     (!p.conv || (p.Cin & 3) == 0);
  const bool w_vec = p.w_kn ? (((p.ldw & 3) == 0) && ((p.N & 3) == 0) && ((((uintptr_t)W) & 15) == 0))
                            : (((p.ldw & 3) == 0) && ((p.K & 3) == 0) && ((((uintptr_t)W) & 15) == 0));

  // --- per-thread A row bookkeeping (rows are fixed across the K loop) ---
  // vector path: slot i -> row = (tid + i*NT) / 4, kq = ((tid + i*NT) % 4) * 4
  // scalar path: element e (4*AV of them) -> idx = tid + e*NT, row = idx / BK, k = idx % BK
  int a_b[AV], a_t[AV], a_v[AV];
  if (p.conv && a_vec) {
#pragma unroll
    for (int i = 0; i < AV; ++i) {
      int r = row0 + (tid + i * NT) / 4;
      int bt = r / p.V;
      a_v[i] = r - bt * p.V;
      a_b[i] = bt / p.T;
      a_t[i] = bt - a_b[i] * p.T;
    }
  }

  float4 ra[AV], rb[BV];

  auto load_a_elem = [&](int r, int k) -> float {
    if (r >= p.M || k >= p.K) return 0.f;
    float v;
    if (p.conv) {
      int bt = r / p.V, vv = r - bt * p.V, b = bt / p.T, t = bt - b * p.T;
      int tap = k / p.Cin, ci = k - tap * p.Cin;
      int ts = reflect_idx(t + tap - p.taps / 2, p.T) / p.tdiv;
      v = A[((long long)(b * (p.T / p.tdiv) + ts) * p.V + vv) * p.lda + ci];
    } else {
      v = A[(long long)r * p.lda + k];
    }
    return p.a_lrelu ? lrelu02(v) : v;
  };

  auto fetch = [&](int kt) {
    const int k0 = kt * BK;
    if (a_vec) {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        int idx = tid + i * NT;
        int r = row0 + idx / 4, k = k0 + (idx % 4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.M && k < p.K) {
          const float* src;
          if (p.conv) {
            int tap = k / p.Cin, ci = k - tap * p.Cin;
            int ts = reflect_idx(a_t[i] + tap - p.taps / 2, p.T) / p.tdiv;
            src = A + ((long long)(a_b[i] * (p.T / p.tdiv) + ts) * p.V + a_v[i]) * p.lda + ci;
          } else {
            src = A + (long long)r * p.lda + k;
          }
          v = __ldg(reinterpret_cast<const float4*>(src));
          if (p.a_lrelu) { v.x = lrelu02(v.x); v.y = lrelu02(v.y); v.z = lrelu02(v.z); v.w = lrelu02(v.w); }
        }
        ra[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        float t4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int idx = tid + (i * 4 + e) * NT;
          t4[e] = load_a_elem(row0 + idx / BK, k0 + idx % BK);
        }
        ra[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
      }
    }
    if (p.w_kn) {
      // W[k][n], contiguous along n
      if (w_vec) {
#pragma unroll
        for (int i = 0; i < BV; ++i) {
          int idx = tid + i * NT;
          int k = k0 + idx / (BN / 4), n = col0 + (idx % (BN / 4)) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < p.K && n < p.N) v = __ldg(reinterpret_cast<const float4*>(W + (long long)k * p.ldw + n));
          rb[i] = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < BV; ++i) {
          float t4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int idx = tid + i * NT;
            int k = k0 + idx / (BN / 4), n = col0 + (idx % (BN / 4)) * 4 + e;
            t4[e] = (k < p.K && n < p.N) ? W[(long long)k * p.ldw + n] : 0.f;
          }
          rb[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
        }
      }
    } else {
      if (w_vec) {
#pragma unroll
        for (int i = 0; i < BV; ++i) {
          int idx = tid + i * NT;
          int n = col0 + idx / 4, k = k0 + (idx % 4) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < p.N && k < p.K) v = __ldg(reinterpret_cast<const float4*>(W + (long long)n * p.ldw + k));
          rb[i] = v;
        }
      } else {
#pragma unroll
        for (int i = 0; i < BV; ++i) {
          float t4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int idx = tid + (i * 4 + e) * NT;
            int n = col0 + idx / BK, k = k0 + idx % BK;
            t4[e] = (n < p.N && k < p.K) ? W[(long long)n * p.ldw + k] : 0.f;
          }
          rb[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
        }
      }
    }
  };

  auto stash = [&](int buf) {
    if (a_vec) {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        int idx = tid + i * NT;
        int r = idx / 4, k = (idx % 4) * 4;
        As[buf][k + 0][r] = ra[i].x;
        As[buf][k + 1][r] = ra[i].y;
        As[buf][k + 2][r] = ra[i].z;
        As[buf][k + 3][r] = ra[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        const float t4[4] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int idx = tid + (i * 4 + e) * NT;
          As[buf][idx % BK][idx / BK] = t4[e];
        }
      }
    }
    if (p.w_kn) {
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        int idx = tid + i * NT;
        int k = idx / (BN / 4), n = (idx % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][n]) = rb[i];
      }
    } else if (w_vec) {
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        int idx = tid + i * NT;
        int n = idx / 4, k = (idx % 4) * 4;
        Bs[buf][k + 0][n] = rb[i].x;
        Bs[buf][k + 1][n] = rb[i].y;
        Bs[buf][k + 2][n] = rb[i].z;
        Bs[buf][k + 3][n] = rb[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        const float t4[4] = {rb[i].x, rb[i].y, rb[i].z, rb[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int idx = tid + (i * 4 + e) * NT;
          Bs[buf][idx % BK][idx / BK] = t4[e];
        }
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) fetch(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      // thread owns rows {ty*4+i} and (TM==8) {BM/2 + ty*4 + i}; same split for columns
      {
        float4 v = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        if (TM == 8) {
          float4 u = *reinterpret_cast<const float4*>(&As[cur][k][BM / 2 + ty * 4]);
          a[TM - 4] = u.x; a[TM - 3] = u.y; a[TM - 2] = u.z; a[TM - 1] = u.w;
        }
      }
      {
        float4 v = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        if (TN == 8) {
          float4 u = *reinterpret_cast<const float4*>(&Bs[cur][k][BN / 2 + tx * 4]);
          b[TN - 4] = u.x; b[TN - 3] = u.y; b[TN - 2] = u.z; b[TN - 1] = u.w;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(cur ^ 1);
    __syncthreads();
  }

  // --- epilogue ---
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = row0 + ((TM == 8 && i >= 4) ? BM / 2 + ty * 4 + (i - 4) : ty * 4 + i);
    if (r >= p.M) continue;
    const float* brow = nullptr;
    if (p.bias) brow = p.bias_period > 0 ? p.bias + (long long)(r % p.bias_period) * p.N : p.bias;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = col0 + ((TN == 8 && j >= 4) ? BN / 2 + tx * 4 + (j - 4) : tx * 4 + j);
      if (c >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (brow) v += brow[c];
      if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == ACT_GELU) v = gelu_erf(v);
      else if (p.act == ACT_LRELU) v = lrelu02(v);
      if (R) v += R[(long long)r * p.ldr + c];
      C[(long long)r * p.ldc + c] = v;
    }
  }
}

template <int BM, int BN, int TM, int TN>
int launch(const GemmParams& p, cudaStream_t s) {
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), p.nz);
  launch_k(sgemm_kernel<BM, BN, TM, TN>, grid, (BM / TM) * (BN / TN), 0, s, p);
  count_launch();
  MOCHA_LAUNCH_CHECK("sgemm_kernel");
  return MOCHA_OK;
}

}  // namespace

int gemm_f32(const GemmParams& p, cudaStream_t stream) {
  MOCHA_CHECK_ARG(p.A && p.W && p.C, "gemm_f32: null operand");
  MOCHA_CHECK_ARG(p.M > 0 && p.N > 0 && p.K > 0, "gemm_f32: bad shape M=%d N=%d K=%d", p.M, p.N, p.K);
  MOCHA_CHECK_ARG(p.nz >= 1 && p.nz2 >= 1 && p.nz % p.nz2 == 0, "gemm_f32: bad batch nz=%d nz2=%d", p.nz, p.nz2);
  if (p.conv) {
    MOCHA_CHECK_ARG(p.T > 0 && p.V > 0 && p.taps > 0 && p.Cin > 0 && p.tdiv > 0, "gemm_f32: bad conv geometry");
    MOCHA_CHECK_ARG(p.K == p.taps * p.Cin, "gemm_f32: conv K=%d != taps*Cin=%d", p.K, p.taps * p.Cin);
    MOCHA_CHECK_ARG(p.Cin % BK == 0, "gemm_f32: conv Cin=%d must be a multiple of %d", p.Cin, BK);
    MOCHA_CHECK_ARG(p.M % (p.T * p.V) == 0, "gemm_f32: conv M=%d not a multiple of T*V", p.M);
    MOCHA_CHECK_ARG(p.T % p.tdiv == 0 && p.taps / 2 < p.T, "gemm_f32: conv T/tdiv/taps mismatch");
  }
  // tile choice: fill 148 SMs; big tiles only when there are enough of them
  const long long t128 = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 128) * p.nz;
  const long long t64 = (long long)ceil_div(p.M, 64) * ceil_div(p.N, 64) * p.nz;
  if (t128 >= 2 * 148) return launch<128, 128, 8, 8>(p, stream);
  if (t64 >= 148) return launch<64, 64, 4, 4>(p, stream);
  return launch<32, 32, 4, 4>(p, stream);
}

}  // namespace mocha
