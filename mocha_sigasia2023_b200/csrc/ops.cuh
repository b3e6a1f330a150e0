// Bandwidth/latency-bound building blocks of the path (graph aggregation, pooling, instance /
// layer normalisation, softmax, CVAE token assembly). All activations are channel-last:
// Generator tensors are [B, T, V, C] (row = (b*T + t)*V + v), token tensors are [B, n, C].
#pragma once
#include "common.cuh"

namespace mocha {

// out[(bt,w), k*C + c] = sum_u f(in[(bt,u), c]) * A[k,u,w]      (f = LeakyReLU(0.2) if lrelu)
// Reference: SpatialConv einsum 'nkctv,kvw->nctw' (net/blocks.py:64) commuted in front of the
// 1x1 convolution, with the pre-activation of STGCN_Block.forward (net/blocks.py:125-129).
// out and/or out16 (bf16 copy for a tensor-core consumer) may be given
int graph_agg_first(const float* in, const float* A, float* out, int BT, int V, int C, int Kk,
                    int lrelu, cudaStream_t s, __nv_bfloat16* out16 = nullptr);

// Fused Conv2d 1x1 (Cin -> C) + LeakyReLU + graph aggregation of the tensor-core path:
// out16[(bt,w), k*C + c] = sum_u lrelu(X[(bt,u), :] . Wemb[c, :] + bemb[c]) * A[k,u,w]
// ldo > Kk*C: rows are padded to ldo columns; columns Kk*C .. Kk*C+Kk-1 hold sum_u A[k,u,w] (for biases folded
// into the following GEMM as extra K columns), the rest zeros
int embed_graph_agg(const float* X, const float* Wemb, const float* bemb, const float* A, __nv_bfloat16* out16, int BT,
                    int V, int Cin, int C, int Kk, cudaStream_t s, int ldo = 0);
// copies the 2*pad reflect-padding frames of xp [B, T + 2*pad, row_elems] from its already written interior
int reflect_border_fill(__nv_bfloat16* xp, int B, int T, int pad, long long row_elems, cudaStream_t s);

// out[(bt,w), c] = sum_k sum_u in[(bt,u), k*C + c] * A2[k,u,w]   (U input nodes, Wn output nodes)
// Same einsum applied after the 1x1 convolution (used by to_mot's JointBlock, where the
// body-part -> joint un-pooling is folded into A2 on the host).
int graph_agg_kv(const float* in, const float* A2, float* out, int BT, int U, int Wn, int C, int Kk,
                 cudaStream_t s);

// Same aggregation emitted as the next temporal conv's bf16 operand: out16 [B, Ts*tdiv + 2*pad, Wn, C] with
// nearest x tdiv up-sampling in time and reflect padding (tensor-core path of to_mot)
int graph_agg_kv_pad16(const float* in, const float* A2, __nv_bfloat16* out16, int B, int Ts, int tdiv, int pad, int U,
                       int Wn, int C, int Kk, cudaStream_t s);

// out[b,t',p,c] = (1/tp) * sum_{dt<tp} sum_v in[b, tp*t'+dt, v, c] * Wp[v,p]
// Reference: PoolJointToBodypart.forward (net/graph.py:463-465) + nn.AvgPool2d((tp,1)) (model.py:47).
int pool_joint_body(const float* in, const float* Wp, float* out, int B, int T, int V, int P, int C,
                    int tp, cudaStream_t s);

// bf16 input variant fused with the next block's LeakyReLU + graph aggregation (tensor-core path):
// out16[(b,t',w), k*C + c] = sum_u lrelu(pool(in)[b,t',u,c]) * A[k,u,w]
int pool_graph_agg(const __nv_bfloat16* in, const float* Wp, const float* A, __nv_bfloat16* out16, int B, int T, int V,
                   int P, int C, int tp, int Kk, cudaStream_t s);

// Instance norm over tokens per (b, channel): unbiased std, eps added to std
// (mean_variance_norm, net/transformer.py:13-20). Optional AdaIN modulation
// y = (1+gamma)*IN(x) + beta with gb = [B, 2C] (gamma | beta) (AdaIN.forward, transformer.py:108-113)
// and optional second output y2 = (y - tab_mean[n,c]) / tab_std[n,c] (test_fullframework.py:293,442).
int instance_norm_tokens(const float* x, int B, int n, int C, float eps, const float* gb, float* y,
                         const float* tab_mean, const float* tab_std, float* y2, cudaStream_t s,
                         __nv_bfloat16* y16 = nullptr, __nv_bfloat16* y2h = nullptr,
                         const float* y2h_center = nullptr);

// Tensor-core path: y = AdaIN(x) (fp32) and q16 = IN(y) (bf16) from ONE pass over x (n <= 128, C % 64 == 0):
// IN(g u + be) = u * g / (|g| std_u + eps) with u = IN(x), std_u = std / (std + eps)
int adain_norm_tokens(const float* x, int B, int n, int C, float eps, const float* gb, float* y, __nv_bfloat16* q16,
                      cudaStream_t s);

// mean over tokens: out[b,c] = mean_n x[b,n,c]  (AdaptiveAvgPool1d(1), transformer.py:102)
int token_mean(const float* x, int B, int n, int C, float* out, cudaStream_t s);
// to_mot's output layer + de-normalisation in one pass: ytil [R, N] = x [R, K] (bf16) W^T [N, K] + b, y = ytil * sd + mu
// with [period, N] tables indexed by row % period
bool out_conv_affine_supported(int K, int N);
int out_conv_affine(const __nv_bfloat16* x, const float* W, const float* bias, const float* mu, const float* sd, float* ytil,
                    float* y, int R, int K, int N, int period, cudaStream_t s);
// last CVAE prior layer on its two read rows (tokens 0 / 1 of x [B, np, D]; kv bf16 [B*np, 2D] from the K/V projection):
// q projection, attention, out-projection + LayerNorm, FFN + LayerNorm in one launch -> out [2B, D] fp32
bool cvae_prior_last_supported(int D, int H, int dff, int np);
int cvae_prior_last(const float* x, int B, int np, const __nv_bfloat16* kv, const __nv_bfloat16* wq, const float* bq,
                    const __nv_bfloat16* wo, const float* bo, const float* g1, const float* be1, const __nv_bfloat16* w1,
                    const float* b1, const __nv_bfloat16* w2, const float* b2, const float* g2, const float* be2, int H, int dff,
                    float eps, float* out, cudaStream_t s);
// AdaIN parameter MLPs of all decoder layers in one launch: gb [nlayers, B, 2D] = W2 lrelu(W1 mean_tokens(cha) + b1) + b2
// (bf16 weights, fp32 activations)
bool style_mlp_supported(int D, int nlayers);
int style_mlp(const float* cha, int B, int n, int D, int nlayers, const __nv_bfloat16* const* w1, const float* const* b1,
              const __nv_bfloat16* const* w2, const float* const* b2, float* gb, cudaStream_t s);

// in-place softmax over the last dim of S[rows, ncols] after multiplying by scale
int softmax_rows(float* S, long long rows, int ncols, float scale, cudaStream_t s);

// y = LayerNorm(x + r) * g + b over the last dim C (post-LN TransformerEncoder/DecoderLayer)
// optional y2 = y * tab_std[row % period, c] + tab_mean[row % period, c]
int add_layernorm(const float* x, const float* r, const float* g, const float* b, float* y, long long rows,
                  int C, float eps, const float* tab_mean, const float* tab_std, int period, float* y2,
                  cudaStream_t s, __nv_bfloat16* y16 = nullptr);

// CVAE token assembly (model_CVAE.py:70-76, :159-164)
int cvae_prior_tokens(const float* mu_token, const float* logvar_token, const float* cond, const float* pe,
                      float* tok, int B, int ncond, int C, cudaStream_t s, __nv_bfloat16* tok16 = nullptr);
// z = mu + eps*exp(0.5*logvar) (eps may be null => z = mu); mem = [z ; cond]
int cvae_memory(const float* prior_out, int prior_tokens, const float* eps, const float* cond, float* mem,
                float* mu_out, float* logvar_out, int B, int ncond, int C, cudaStream_t s,
                __nv_bfloat16* mem16 = nullptr);
// cond = [ (src_cnt - m0)/s0 ; (prev - m1)/s1 ]  (test_fullframework.py:446-447)
int cvae_condition(const float* src_cnt, const float* prev, const float* m0, const float* s0, const float* m1,
                   const float* s1, float* cond, int B, int n, int C, cudaStream_t s);
// out[r, c] = x[r, c] * sd[(r % period), c] + mu[(r % period), c]
// x may be a padded view (row pitch ld_in >= C); `copy` optionally receives the dense [rows, C] copy of x
int affine_rows(const float* x, const float* mu, const float* sd, float* out, long long rows, int C, int period,
                cudaStream_t s, int ld_in = 0, float* copy = nullptr);
// out[b, n, c] = x[n, c]  (broadcast a table over the batch)
int broadcast_rows(const float* x, float* out, int B, long long n_elems, cudaStream_t s,
                   __nv_bfloat16* out16 = nullptr);
// out = a + table[(r % period)]
int add_table(const float* a, const float* table, float* out, long long rows, int C, int period, cudaStream_t s);

// out[b, i, :] = x[b, i, :] for i < take; fp32 and/or bf16 outputs (x16 optional bf16 source)
int gather_token_rows(const float* x, const __nv_bfloat16* x16, int n, int take, int C, int B, float* out,
                      __nv_bfloat16* out16, cudaStream_t s);

int cast_f32_bf16(const float* x, __nv_bfloat16* y, long long n, cudaStream_t s);

}  // namespace mocha
