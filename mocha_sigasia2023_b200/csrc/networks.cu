// Stage-level orchestration of the Generator and CVAE forward passes behind the C ABI
// (include/mocha_b200.h). Each stage is a fixed sequence of kernel launches on the caller's stream
// over caller-provided workspace: no allocation, no synchronisation, CUDA-graph capturable.
#include "../../include/mocha_b200.h"
#include "common.cuh"
#include "gemm_f32.cuh"
#include "gemm_tc.cuh"
#include "fused.cuh"
#include "networks_bf16.cuh"
#include "ops.cuh"

using namespace mocha;

namespace {

struct Ctx {
  cudaStream_t s;
  int precision;
  Workspace* ws;
};

// C[M,N] = act(prologue(A) W^T + bias) (+res); fp32 FFMA, tcgen05 bf16 or tcgen05 3xTF32 depending on ctx.precision.
// Shapes the tensor-core kernel does not take (tiny N/K, gathers) stay on the fp32 kernel.
int dense(const Ctx& c, const float* A, int lda, const float* W, const float* bias, int bias_period,
          const float* res, float* C, int M, int N, int K, int act, int a_lrelu = 0) {
  if (c.precision == MOCHA_BF16 && tc_linear_supported(M, N, K) && lda == K) {
    return tc_linear(A, W, bias, bias_period, res, C, M, N, K, act, a_lrelu, *c.ws, c.s);
  }
  if (c.precision == MOCHA_TF32X3 && tc_linear_tf32x3_supported(M, N, K) && lda == K)
    return tc_linear_tf32x3(A, W, bias, bias_period, res, C, M, N, K, act, a_lrelu, *c.ws, c.s);
  GemmParams p;
  p.A = A; p.W = W; p.C = C;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldw = K; p.ldc = N;
  p.bias = bias; p.bias_period = bias_period;
  p.act = act; p.res = res; p.ldr = N; p.a_lrelu = a_lrelu;
  return gemm_f32(p, c.s);
}

// reflect-padded temporal convolution as implicit GEMM (blocks.py:113-118,:132)
int tconv(const Ctx& c, const float* A, const float* W, const float* bias, int bias_period, float* C, int B,
          int T, int V, int Cin, int Cout, int taps, int tdiv) {
  if (c.precision == MOCHA_BF16 && tc_tconv_supported(B, T, V, Cin, Cout, taps))
    return tc_tconv(A, W, bias, bias_period, C, B, T, V, Cin, Cout, taps, tdiv, *c.ws, c.s);
  if (c.precision == MOCHA_TF32X3 && tc_tconv_tf32x3_supported(B, T, V, Cin, Cout, taps))
    return tc_tconv_tf32x3(A, W, bias, bias_period, C, B, T, V, Cin, Cout, taps, tdiv, *c.ws, c.s);
  GemmParams p;
  p.A = A; p.W = W; p.C = C;
  p.M = B * T * V; p.N = Cout; p.K = taps * Cin;
  p.lda = Cin; p.ldw = taps * Cin; p.ldc = Cout;
  p.conv = 1; p.T = T; p.V = V; p.taps = taps; p.Cin = Cin; p.tdiv = tdiv;
  p.bias = bias; p.bias_period = bias_period;
  return gemm_f32(p, c.s);
}

// softmax(Q K^T * scale) V for B*H (batch, head) problems; operands are strided views into fused
// projection buffers. S is a [B,H,nq,nkv] scratch.
int attention(const Ctx& c, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int B,
              int H, int nq, int nkv, int dh, float* S, float* out, int ldo) {
  if (c.precision == MOCHA_BF16 && tc_attention_supported(nq, nkv, dh))
    return tc_attention(q, ldq, k, ldk, v, ldv, B, H, nq, nkv, dh, S, out, ldo, *c.ws, c.s);
  if (c.precision == MOCHA_TF32X3 && tc_attention_tf32x3_supported(nq, nkv, dh) && nq >= 16) {
    // (tiny query counts - the cached decoder-0 table - stay on the fp32 kernel)
    const int rc = tc_attention_tf32x3(q, ldq, k, ldk, v, ldv, B, H, nq, nkv, dh, S, out, ldo, *c.ws, c.s);
    if (rc != MOCHA_ERR_WORKSPACE) return rc;   // without room for the split operands the FFMA products below still apply
  }
  GemmParams p;
  p.A = q; p.lda = ldq; p.sA1 = (long long)nq * ldq; p.sA2 = dh;
  p.W = k; p.ldw = ldk; p.sW1 = (long long)nkv * ldk; p.sW2 = dh;
  p.C = S; p.ldc = nkv; p.sC1 = (long long)H * nq * nkv; p.sC2 = (long long)nq * nkv;
  p.M = nq; p.N = nkv; p.K = dh; p.nz = B * H; p.nz2 = H;
  MOCHA_TRY(gemm_f32(p, c.s));
  MOCHA_TRY(softmax_rows(S, (long long)B * H * nq, nkv, 1.0f / sqrtf((float)dh), c.s));
  GemmParams g;
  g.A = S; g.lda = nkv; g.sA1 = (long long)H * nq * nkv; g.sA2 = (long long)nq * nkv;
  g.W = v; g.w_kn = 1; g.ldw = ldv; g.sW1 = (long long)nkv * ldv; g.sW2 = dh;
  g.C = out; g.ldc = ldo; g.sC1 = (long long)nq * ldo; g.sC2 = dh;
  g.M = nq; g.N = dh; g.K = nkv; g.nz = B * H; g.nz2 = H;
  return gemm_f32(g, c.s);
}

int check_dims(const mocha_dims& d) {
  MOCHA_CHECK_ARG(d.T > 0 && d.V > 0 && d.Cin > 0 && d.C0 > 0 && d.D > 0 && d.P > 0 && d.tp > 0, "dims: non-positive");
  MOCHA_CHECK_ARG(d.T % d.tp == 0, "dims: T %% tp != 0");
  MOCHA_CHECK_ARG(d.C0 % 16 == 0 && d.D % 16 == 0, "dims: C0 and D must be multiples of 16");
  MOCHA_CHECK_ARG(d.enc_depth >= 1 && d.enc_depth <= MOCHA_MAX_DEPTH && d.dec_depth >= 1 && d.dec_depth <= MOCHA_MAX_DEPTH,
                  "dims: depth out of range");
  MOCHA_CHECK_ARG(d.heads > 0 && d.enc_dh > 0 && d.dec_dh > 0 && d.mlp > 0, "dims: bad transformer geometry");
  MOCHA_CHECK_ARG((d.T / d.tp) * d.P <= 256, "dims: more than 256 tokens");
  return MOCHA_OK;
}

inline int ntok(const mocha_dims& d) { return (d.T / d.tp) * d.P; }

#define WS_GUARD(ws, name)                                                                         \
  do {                                                                                             \
    if ((ws).overflow)                                                                             \
      return set_error(MOCHA_ERR_WORKSPACE, "%s: workspace too small (%zu B given, %zu B needed)", \
                       name, (ws).cap, (ws).off);                                                  \
  } while (0)

size_t pad256(size_t n) { return align_up(n, 256); }

}  // namespace

// ------------------------------------------------------------------------------------------------
// mot_embedding
// ------------------------------------------------------------------------------------------------
extern "C" size_t mocha_embed_workspace_bytes(const mocha_dims* d, int B) {
  if (!d || B <= 0) return 0;
  const size_t R = (size_t)B * d->T * d->V, R2 = (size_t)B * ntok(*d);
  size_t n = 0;
  n += pad256(R * d->C0 * 4);               // h0
  n += pad256(R * d->Kj * d->C0 * 4);       // agg
  n += pad256(R * d->D * 4) * 2;            // g, h1
  n += pad256(R2 * d->D * 4);               // pooled
  n += pad256(R2 * d->Kb * d->D * 4);       // agg2
  n += pad256(R2 * d->D * 4);               // g2
  n += tc_scratch_bytes(R, d->Kj * d->C0 > d->D ? d->Kj * d->C0 : d->D);
  n += tc_tconv_scratch_bytes(B, d->T, d->V, d->D, d->taps_j);
  return n + 4096;
}

extern "C" int mocha_embed_fwd(const mocha_generator_weights* w, const float* X, int B, float* tokens,
                               int add_pos_emb, int precision, void* workspace, size_t workspace_bytes,
                               mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && X && tokens && B > 0, "mocha_embed_fwd: null/empty argument");
  MOCHA_TRY(check_dims(w->dims));
  const mocha_dims& d = w->dims;
  Workspace ws(workspace, workspace_bytes);
  if (precision == MOCHA_BF16 && bf16_path_supported(d))
    return embed_bf16(w, X, B, tokens, add_pos_emb, ws, (cudaStream_t)stream);
  Ctx c{(cudaStream_t)stream, precision, &ws};
  const int R = B * d.T * d.V, Tp = d.T / d.tp, R2 = B * Tp * d.P;
  float* h0 = ws.take<float>((size_t)R * d.C0);
  float* agg = ws.take<float>((size_t)R * d.Kj * d.C0);
  float* g = ws.take<float>((size_t)R * d.D);
  float* h1 = ws.take<float>((size_t)R * d.D);
  float* pooled = ws.take<float>((size_t)R2 * d.D);
  float* agg2 = ws.take<float>((size_t)R2 * d.Kb * d.D);
  float* g2 = ws.take<float>((size_t)R2 * d.D);
  WS_GUARD(ws, "mocha_embed_fwd");

  // Conv2d 1x1 Cin->C0 (model.py:44)
  MOCHA_TRY(dense(c, X, d.Cin, w->emb_w, w->emb_b, 0, nullptr, h0, R, d.C0, d.Cin, ACT_NONE));
  // JointBlock: LeakyReLU -> graph aggregation -> 1x1 conv (K = Kj*C0) -> temporal conv (model.py:109-134)
  MOCHA_TRY(graph_agg_first(h0, w->A_j, agg, B * d.T, d.V, d.C0, d.Kj, 1, c.s));
  MOCHA_TRY(dense(c, agg, d.Kj * d.C0, w->jb_gcn_w, w->jb_gcn_bias2d, d.V, nullptr, g, R, d.D, d.Kj * d.C0, ACT_NONE));
  MOCHA_TRY(tconv(c, g, w->jb_tcn_w, w->jb_tcn_b, 0, h1, B, d.T, d.V, d.D, d.D, d.taps_j, 1));
  // PoolJointToBodypart + AvgPool2d((tp,1)) (model.py:46-47)
  MOCHA_TRY(pool_joint_body(h1, w->pool_w, pooled, B, d.T, d.V, d.P, d.D, d.tp, c.s));
  // BodyBlock (model.py:137-162)
  MOCHA_TRY(graph_agg_first(pooled, w->A_b, agg2, B * Tp, d.P, d.D, d.Kb, 1, c.s));
  MOCHA_TRY(dense(c, agg2, d.Kb * d.D, w->bb_gcn_w, w->bb_gcn_bias2d, d.P, nullptr, g2, R2, d.D, d.Kb * d.D, ACT_NONE));
  if (add_pos_emb)
    MOCHA_TRY(tconv(c, g2, w->bb_tcn_w, w->tok_bias_pos, Tp * d.P, tokens, B, Tp, d.P, d.D, d.D, d.taps_b, 1));
  else
    MOCHA_TRY(tconv(c, g2, w->bb_tcn_w, w->bb_tcn_b, 0, tokens, B, Tp, d.P, d.D, d.D, d.taps_b, 1));
  return MOCHA_OK;
}

// ------------------------------------------------------------------------------------------------
// encoder: x = Attn(x,x)+x ; x = FF(x)+x, no normalisation (transformer.py:90-95 with adain=False)
// ------------------------------------------------------------------------------------------------
extern "C" size_t mocha_encoder_workspace_bytes(const mocha_dims* d, int B) {
  if (!d || B <= 0) return 0;
  const size_t n = ntok(*d), R = (size_t)B * n, inner = (size_t)d->heads * d->enc_dh;
  size_t bytes = 0;
  bytes += pad256(R * 3 * inner * 4);                  // qkv
  bytes += pad256((size_t)B * d->heads * n * n * 4);   // scores
  bytes += pad256(R * inner * 4);                      // attention output
  bytes += pad256(R * d->D * 4) * 2;                   // x ping-pong
  bytes += pad256(R * d->mlp * 4);                     // hidden
  bytes += tc_scratch_bytes(R, inner > (size_t)d->mlp ? inner : d->mlp);
  bytes += tc_attention_scratch_bytes(B, d->heads, (int)n, (int)n, d->enc_dh);
  return bytes + 4096;
}

extern "C" int mocha_encoder_fwd(const mocha_generator_weights* w, const float* tokens, int B, float* encoded,
                                 int precision, void* workspace, size_t workspace_bytes, mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && tokens && encoded && B > 0, "mocha_encoder_fwd: null/empty argument");
  MOCHA_TRY(check_dims(w->dims));
  const mocha_dims& d = w->dims;
  Workspace ws(workspace, workspace_bytes);
  if (precision == MOCHA_BF16 && bf16_path_supported(d))
    return encoder_bf16(w, tokens, B, encoded, ws, (cudaStream_t)stream);
  Ctx c{(cudaStream_t)stream, precision, &ws};
  const int n = ntok(d), R = B * n, inner = d.heads * d.enc_dh;
  float* qkv = ws.take<float>((size_t)R * 3 * inner);
  float* S = ws.take<float>((size_t)B * d.heads * n * n);
  float* att = ws.take<float>((size_t)R * inner);
  float* xa = ws.take<float>((size_t)R * d.D);
  float* xb = ws.take<float>((size_t)R * d.D);
  float* hid = ws.take<float>((size_t)R * d.mlp);
  WS_GUARD(ws, "mocha_encoder_fwd");

  const float* x = tokens;
  for (int l = 0; l < d.enc_depth; ++l) {
    const mocha_enc_layer& L = w->enc[l];
    MOCHA_CHECK_ARG(L.wqkv && L.wo && L.bo && L.w1 && L.b1 && L.w2 && L.b2, "mocha_encoder_fwd: layer %d weights missing", l);
    MOCHA_TRY(dense(c, x, d.D, L.wqkv, nullptr, 0, nullptr, qkv, R, 3 * inner, d.D, ACT_NONE));
    MOCHA_TRY(attention(c, qkv, 3 * inner, qkv + inner, 3 * inner, qkv + 2 * inner, 3 * inner, B, d.heads, n, n,
                        d.enc_dh, S, att, inner));
    MOCHA_TRY(dense(c, att, inner, L.wo, L.bo, 0, x, xa, R, d.D, inner, ACT_NONE));
    MOCHA_TRY(dense(c, xa, d.D, L.w1, L.b1, 0, nullptr, hid, R, d.mlp, d.D, ACT_GELU));
    float* dst = (l == d.enc_depth - 1) ? encoded : xb;
    MOCHA_TRY(dense(c, hid, d.mlp, L.w2, L.b2, 0, xa, dst, R, d.D, d.mlp, ACT_NONE));
    x = dst;
  }
  return MOCHA_OK;
}

extern "C" int mocha_cnt_features(const float* x, int B, int n, int C, float eps, float* cnt,
                                  const float* cnt_mean, const float* cnt_std, float* cnt_nm, void* cnt_nm16,
                                  const float* cnt_nm16_center, mocha_stream_t stream) {
  return instance_norm_tokens(x, B, n, C, eps, nullptr, cnt, cnt_mean, cnt_std, cnt_nm, (cudaStream_t)stream, nullptr,
                              static_cast<__nv_bfloat16*>(cnt_nm16), cnt_nm16_center);
}

// ------------------------------------------------------------------------------------------------
// decoder: per layer x = AdaIN(x, sty); x = Attn(q=IN(x), k=IN(sty), v=sty) + x; x = FF(x) + x
// (transformer.py:90-113)
// ------------------------------------------------------------------------------------------------
extern "C" size_t mocha_decoder_workspace_bytes(const mocha_dims* d, int B) {
  if (!d || B <= 0) return 0;
  const size_t n = ntok(*d), R = (size_t)B * n, inner = (size_t)d->heads * d->dec_dh;
  size_t bytes = 0;
  bytes += pad256((size_t)B * d->D * 4);           // style mean
  bytes += pad256((size_t)B * 2 * d->D * 4) * (1 + MOCHA_MAX_DEPTH);   // style hidden, gamma|beta of every layer
  bytes += pad256(R * d->D * 4) * 5;               // sty_in, x1, qin, x2, xb
  bytes += pad256(R * inner * 4) * 4;              // q, k, v, att
  bytes += pad256(R * inner * 2) * 2 * (MOCHA_MAX_DEPTH - 1);   // k | v of the later layers, projected ahead (bf16 path)
  bytes += pad256((size_t)B * d->heads * n * n * 4);
  bytes += pad256(R * d->mlp * 4);
  bytes += tc_scratch_bytes(R, inner > (size_t)d->mlp ? inner : d->mlp);
  bytes += tc_attention_scratch_bytes(B, d->heads, (int)n, (int)n, d->dec_dh);
  return bytes + 4096;
}

extern "C" int mocha_decoder_fwd(const mocha_generator_weights* w, const float* src, const float* cha, int B,
                                 float* decoded, int precision, void* workspace, size_t workspace_bytes,
                                 mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && src && cha && decoded && B > 0, "mocha_decoder_fwd: null/empty argument");
  MOCHA_TRY(check_dims(w->dims));
  const mocha_dims& d = w->dims;
  Workspace ws(workspace, workspace_bytes);
  if (precision == MOCHA_BF16 && bf16_path_supported(d))
    return decoder_bf16(w, src, cha, B, decoded, ws, (cudaStream_t)stream);
  Ctx c{(cudaStream_t)stream, precision, &ws};
  const int n = ntok(d), R = B * n, inner = d.heads * d.dec_dh;
  const float eps = 1e-5f;
  float* smean = ws.take<float>((size_t)B * d.D);
  float* shid = ws.take<float>((size_t)B * 2 * d.D);
  float* gb = ws.take<float>((size_t)B * 2 * d.D);
  float* sty_in = ws.take<float>((size_t)R * d.D);
  float* x1 = ws.take<float>((size_t)R * d.D);
  float* qin = ws.take<float>((size_t)R * d.D);
  float* x2 = ws.take<float>((size_t)R * d.D);
  float* xb = ws.take<float>((size_t)R * d.D);
  float* q = ws.take<float>((size_t)R * inner);
  float* k = ws.take<float>((size_t)R * inner);
  float* v = ws.take<float>((size_t)R * inner);
  float* att = ws.take<float>((size_t)R * inner);
  float* S = ws.take<float>((size_t)B * d.heads * n * n);
  float* hid = ws.take<float>((size_t)R * d.mlp);
  WS_GUARD(ws, "mocha_decoder_fwd");

  // layer-independent functions of the style tokens
  MOCHA_TRY(token_mean(cha, B, n, d.D, smean, c.s));
  MOCHA_TRY(instance_norm_tokens(cha, B, n, d.D, eps, nullptr, sty_in, nullptr, nullptr, nullptr, c.s));

  const float* x = src;
  for (int l = 0; l < d.dec_depth; ++l) {
    const mocha_dec_layer& L = w->dec[l];
    MOCHA_CHECK_ARG(L.sw1 && L.sw2 && L.wq && L.wk && L.wv && L.wo && L.w1 && L.w2, "mocha_decoder_fwd: layer %d weights missing", l);
    // AdaIN parameters: Linear -> LeakyReLU -> Linear on the token-mean of the style
    MOCHA_TRY(dense(c, smean, d.D, L.sw1, L.sb1, 0, nullptr, shid, B, 2 * d.D, d.D, ACT_LRELU));
    MOCHA_TRY(dense(c, shid, 2 * d.D, L.sw2, L.sb2, 0, nullptr, gb, B, 2 * d.D, 2 * d.D, ACT_NONE));
    MOCHA_TRY(instance_norm_tokens(x, B, n, d.D, eps, gb, x1, nullptr, nullptr, nullptr, c.s));
    // Attention with instance-normed q/k inputs (mapping_function, transformer.py:49-55)
    MOCHA_TRY(instance_norm_tokens(x1, B, n, d.D, eps, nullptr, qin, nullptr, nullptr, nullptr, c.s));
    MOCHA_TRY(dense(c, qin, d.D, L.wq, nullptr, 0, nullptr, q, R, inner, d.D, ACT_NONE));
    MOCHA_TRY(dense(c, sty_in, d.D, L.wk, nullptr, 0, nullptr, k, R, inner, d.D, ACT_NONE));
    MOCHA_TRY(dense(c, cha, d.D, L.wv, nullptr, 0, nullptr, v, R, inner, d.D, ACT_NONE));
    MOCHA_TRY(attention(c, q, inner, k, inner, v, inner, B, d.heads, n, n, d.dec_dh, S, att, inner));
    MOCHA_TRY(dense(c, att, inner, L.wo, L.bo, 0, x1, x2, R, d.D, inner, ACT_NONE));
    MOCHA_TRY(dense(c, x2, d.D, L.w1, L.b1, 0, nullptr, hid, R, d.mlp, d.D, ACT_GELU));
    float* dst = (l == d.dec_depth - 1) ? decoded : xb;
    MOCHA_TRY(dense(c, hid, d.mlp, L.w2, L.b2, 0, x2, dst, R, d.D, d.mlp, ACT_NONE));
    x = dst;
  }
  return MOCHA_OK;
}

// ------------------------------------------------------------------------------------------------
// to_mot (model.py:71-80). Nearest x tp up-sampling and the 0/1 un-pooling are pure row copies, so
// the 1x1 convolution of the JointBlock runs on the n_tok body-part rows and the copies are folded
// into the adjacency (tm_A2) and into the temporal-conv gather (tdiv = tp).
// ------------------------------------------------------------------------------------------------
extern "C" size_t mocha_to_mot_workspace_bytes(const mocha_dims* d, int B) {
  if (!d || B <= 0) return 0;
  const size_t R2 = (size_t)B * ntok(*d), R = (size_t)B * d->T * d->V;
  size_t bytes = 0;
  bytes += pad256(R2 * d->Kb * d->D * 4);                 // agg
  bytes += pad256(R2 * d->D * 4) * 2;                     // y1, y2
  bytes += pad256(R2 * d->Kj * d->C0 * 4);                // y3
  bytes += pad256((size_t)B * (d->T / d->tp) * d->V * d->C0 * 4);  // g' (fp32 path)
  bytes += pad256((size_t)B * (d->T + d->taps_j) * d->V * d->C0 * 2);  // g' as the padded bf16 conv operand (bf16 path)
  bytes += pad256(R * d->C0 * 4);                         // y4
  bytes += pad256(R * ((d->Cin + 7) / 8 * 8) * 4);        // Ytil (rows padded to 8 floats on the bf16 path)
  bytes += tc_scratch_bytes(R2, d->Kb * d->D);
  bytes += tc_tconv_scratch_bytes(B, d->T / d->tp, d->P, d->D, d->taps_b);
  bytes += tc_tconv_scratch_bytes(B, d->T, d->V, d->C0, d->taps_j);
  bytes += tc_scratch_bytes(R, d->C0);
  return bytes + 4096;
}

extern "C" int mocha_to_mot_fwd(const mocha_generator_weights* w, const float* tokens, int B, float* Ytil,
                                const float* Y_mean, const float* Y_std, float* Y, int precision,
                                void* workspace, size_t workspace_bytes, mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && tokens && B > 0 && (Ytil || Y), "mocha_to_mot_fwd: null/empty argument");
  MOCHA_CHECK_ARG(!Y || (Y_mean && Y_std), "mocha_to_mot_fwd: Y needs Y_mean/Y_std");
  MOCHA_TRY(check_dims(w->dims));
  const mocha_dims& d = w->dims;
  Workspace ws(workspace, workspace_bytes);
  if (precision == MOCHA_BF16 && bf16_path_supported(d))
    return to_mot_bf16(w, tokens, B, Ytil, Y_mean, Y_std, Y, ws, (cudaStream_t)stream);
  Ctx c{(cudaStream_t)stream, precision, &ws};
  const int Tp = d.T / d.tp, R2 = B * Tp * d.P, R = B * d.T * d.V;
  float* agg = ws.take<float>((size_t)R2 * d.Kb * d.D);
  float* y1 = ws.take<float>((size_t)R2 * d.D);
  float* y2 = ws.take<float>((size_t)R2 * d.D);
  float* y3 = ws.take<float>((size_t)R2 * d.Kj * d.C0);
  float* gp = ws.take<float>((size_t)B * Tp * d.V * d.C0);
  float* y4 = ws.take<float>((size_t)R * d.C0);
  float* ytil_ws = Ytil ? nullptr : ws.take<float>((size_t)R * d.Cin);
  WS_GUARD(ws, "mocha_to_mot_fwd");
  float* yt = Ytil ? Ytil : ytil_ws;

  // BodyBlock(D -> D)
  MOCHA_TRY(graph_agg_first(tokens, w->tm_A_b, agg, B * Tp, d.P, d.D, d.Kb, 1, c.s));
  MOCHA_TRY(dense(c, agg, d.Kb * d.D, w->tm_bb_gcn_w, w->tm_bb_gcn_bias2d, d.P, nullptr, y1, R2, d.D, d.Kb * d.D, ACT_NONE));
  MOCHA_TRY(tconv(c, y1, w->tm_bb_tcn_w, w->tm_bb_tcn_b, 0, y2, B, Tp, d.P, d.D, d.D, d.taps_b, 1));
  // JointBlock(D -> C0): LeakyReLU -> 1x1 conv (on body-part rows) -> adjacency (with un-pool folded)
  MOCHA_TRY(dense(c, y2, d.D, w->tm_jb_gcn_w, w->tm_jb_gcn_b, 0, nullptr, y3, R2, d.Kj * d.C0, d.D, ACT_NONE, 1));
  MOCHA_TRY(graph_agg_kv(y3, w->tm_A2, gp, B * Tp, d.P, d.V, d.C0, d.Kj, c.s));
  // temporal conv over the up-sampled T frames, gathering from the Tp-frame tensor
  MOCHA_TRY(tconv(c, gp, w->tm_jb_tcn_w, w->tm_jb_tcn_b, 0, y4, B, d.T, d.V, d.C0, d.C0, d.taps_j, d.tp));
  // LeakyReLU -> Conv2d 1x1 C0 -> Cin (model.py:77-78)
  MOCHA_TRY(dense(c, y4, d.C0, w->tm_out_w, w->tm_out_b, 0, nullptr, yt, R, d.Cin, d.C0, ACT_NONE, 1));
  if (Y) MOCHA_TRY(affine_rows(yt, Y_mean, Y_std, Y, R, d.Cin, d.V, c.s));
  return MOCHA_OK;
}

// ------------------------------------------------------------------------------------------------
// CVAE.sample (model_CVAE.py:44-46): PriorNet.encode (:70-79) -> reparameterize (:81-87) -> Decoder
// (:159-165). nn.TransformerEncoderLayer / DecoderLayer are post-LN with ReLU (norm_first=False).
// ------------------------------------------------------------------------------------------------
extern "C" size_t mocha_cvae_workspace_bytes(const mocha_cvae_weights* w, int B, int ncond) {
  if (!w || B <= 0 || ncond <= 0) return 0;
  const size_t D = w->D, np = ncond + 2, nm = ncond + 1, nq = w->out_seq;
  const size_t Rp = (size_t)B * np, Rm = (size_t)B * nm, Rq = (size_t)B * nq;
  // qkv / scores / att / proj / hidden are shared by the prior (Rp rows) and the decoder (Rq rows, nq x nq and
  // nq x nm scores): size them for whichever is larger (out_seq may exceed ncond + 2 through the drop-in)
  const size_t Rx = Rp > Rq ? Rp : Rq;
  size_t sc = np * np;
  if (nq * nq > sc) sc = nq * nq;
  if (nq * nm > sc) sc = nq * nm;
  size_t bytes = 0;
  bytes += pad256(Rp * D * 4) * 3;         // tok, xa, xb
  bytes += pad256(Rx * 3 * D * 4);         // qkv (also reused by the decoder)
  bytes += pad256((size_t)B * w->heads * sc * 4);  // scores
  bytes += pad256(Rx * D * 4) * 2;         // att, proj
  bytes += pad256(Rx * w->dff * 4);        // hidden
  bytes += pad256(Rm * D * 4);             // memory
  bytes += pad256(Rm * 2 * D * 4);         // memory K|V
  bytes += pad256(Rm * 2 * D * 2) * (MOCHA_MAX_DEPTH - 1);   // ... of the later decoder layers, projected ahead (bf16 path)
  bytes += pad256(Rq * D * 4) * 3;         // decoder x ping-pong + q
  bytes += tc_scratch_bytes(Rx, w->dff);
  bytes += tc_attention_scratch_bytes(B, w->heads, (int)np, (int)np, (int)(D / w->heads));
  return bytes + 4096;
}

extern "C" int mocha_cvae_sample(const mocha_cvae_weights* w, const float* cond, int B, int ncond,
                                 const float* eps, float* out, float* mu, float* logvar,
                                 const float* out_mean, const float* out_std, float* out_denorm, int precision,
                                 void* workspace, size_t workspace_bytes, mocha_stream_t stream) {
  // out == out_denorm == NULL with mu / logvar given: only the token network runs (CVAE.prior / CVAE.encode)
  MOCHA_CHECK_ARG(w && cond && B > 0 && ncond > 0 && (out || out_denorm || (mu && logvar)), "mocha_cvae_sample: null/empty argument");
  MOCHA_CHECK_ARG(w->D > 0 && w->heads > 0 && w->D % w->heads == 0 && w->dff > 0 && w->out_seq > 0,
                  "mocha_cvae_sample: bad geometry");
  MOCHA_CHECK_ARG(w->depth >= 1 && w->depth <= MOCHA_MAX_DEPTH, "mocha_cvae_sample: depth out of range");
  MOCHA_CHECK_ARG(ncond + 2 <= 512, "mocha_cvae_sample: ncond=%d too long", ncond);
  MOCHA_CHECK_ARG(!out_denorm || (out_mean && out_std), "mocha_cvae_sample: denorm needs its tables");
  Workspace ws(workspace, workspace_bytes);
  {
    const int dh_ = w->D / w->heads;
    if (precision == MOCHA_BF16 && (out || out_denorm) && ncond + 2 <= 256 && w->D % 64 == 0 && w->dff % 64 == 0 && w->out_seq <= ncond + 2 &&
        tc_attention_supported(ncond + 2, ncond + 2, dh_) && tc_attention_supported(w->out_seq, ncond + 1, dh_))
      return cvae_bf16(w, cond, B, ncond, eps, out, mu, logvar, out_mean, out_std, out_denorm, ws, (cudaStream_t)stream);
  }
  Ctx c{(cudaStream_t)stream, precision, &ws};
  const int D = w->D, H = w->heads, dh = D / H, np = ncond + 2, nm = ncond + 1, nq = w->out_seq;
  const int Rp = B * np, Rm = B * nm, Rq = B * nq;
  float* tok = ws.take<float>((size_t)Rp * D);
  float* xa = ws.take<float>((size_t)Rp * D);
  float* xb = ws.take<float>((size_t)Rp * D);
  const int Rx = Rp > Rq ? Rp : Rq;   // buffers shared by the prior (Rp rows) and the decoder (Rq rows)
  size_t sc = (size_t)np * np;
  if ((size_t)nq * nq > sc) sc = (size_t)nq * nq;
  if ((size_t)nq * nm > sc) sc = (size_t)nq * nm;
  float* qkv = ws.take<float>((size_t)Rx * 3 * D);
  float* S = ws.take<float>((size_t)B * H * sc);
  float* att = ws.take<float>((size_t)Rx * D);
  float* proj = ws.take<float>((size_t)Rx * D);
  float* hid = ws.take<float>((size_t)Rx * w->dff);
  float* mem = ws.take<float>((size_t)Rm * D);
  float* memkv = ws.take<float>((size_t)Rm * 2 * D);
  float* da = ws.take<float>((size_t)Rq * D);
  float* db = ws.take<float>((size_t)Rq * D);
  float* dq = ws.take<float>((size_t)Rq * D);
  WS_GUARD(ws, "mocha_cvae_sample");

  // ---- prior network ----
  MOCHA_TRY(cvae_prior_tokens(w->mu_token, w->logvar_token, cond, w->pe, tok, B, ncond, D, c.s));
  float* x = tok;
  float* other = xa;
  for (int l = 0; l < w->depth; ++l) {
    const mocha_cvae_enc_layer& L = w->prior[l];
    MOCHA_CHECK_ARG(L.in_w && L.in_b && L.out_w && L.out_b && L.l1_w && L.l2_w && L.n1_g && L.n2_g,
                    "mocha_cvae_sample: prior layer %d weights missing", l);
    MOCHA_TRY(dense(c, x, D, L.in_w, L.in_b, 0, nullptr, qkv, Rp, 3 * D, D, ACT_NONE));
    MOCHA_TRY(attention(c, qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, B, H, np, np, dh, S, att, D));
    MOCHA_TRY(dense(c, att, D, L.out_w, L.out_b, 0, nullptr, proj, Rp, D, D, ACT_NONE));
    MOCHA_TRY(add_layernorm(x, proj, L.n1_g, L.n1_b, other, Rp, D, w->ln_eps, nullptr, nullptr, 0, nullptr, c.s));
    MOCHA_TRY(dense(c, other, D, L.l1_w, L.l1_b, 0, nullptr, hid, Rp, w->dff, D, ACT_RELU));
    MOCHA_TRY(dense(c, hid, w->dff, L.l2_w, L.l2_b, 0, nullptr, proj, Rp, D, w->dff, ACT_NONE));
    float* nxt = (x == tok) ? xb : x;
    MOCHA_TRY(add_layernorm(other, proj, L.n2_g, L.n2_b, nxt, Rp, D, w->ln_eps, nullptr, nullptr, 0, nullptr, c.s));
    x = nxt;
    other = (x == xa) ? xb : xa;
  }
  // ---- reparameterise, assemble decoder memory [z ; cond] ----
  MOCHA_TRY(cvae_memory(x, np, eps, cond, mem, mu, logvar, B, ncond, D, c.s));
  if (!out && !out_denorm) return MOCHA_OK;   // mu / logvar only

  // ---- decoder ----
  if (!w->dec0_sa) MOCHA_TRY(broadcast_rows(w->pe, da, B, (long long)nq * D, c.s));  // tgt = zeros + pe[:out_seq]
  float* dx = da;
  float* dy = db;
  for (int l = 0; l < w->depth; ++l) {
    const mocha_cvae_dec_layer& L = w->dec[l];
    MOCHA_CHECK_ARG(L.sa_in_w && L.sa_out_w && L.ca_in_w && L.ca_out_w && L.l1_w && L.l2_w && L.n1_g && L.n2_g && L.n3_g,
                    "mocha_cvae_sample: decoder layer %d weights missing", l);
    // self-attention block (layer 0 acts on the constant query: use the cached table when provided)
    if (l == 0 && w->dec0_sa) {
      MOCHA_TRY(broadcast_rows(w->dec0_sa, dy, B, (long long)nq * D, c.s));
    } else {
      MOCHA_TRY(dense(c, dx, D, L.sa_in_w, L.sa_in_b, 0, nullptr, qkv, Rq, 3 * D, D, ACT_NONE));
      MOCHA_TRY(attention(c, qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, B, H, nq, nq, dh, S, att, D));
      MOCHA_TRY(dense(c, att, D, L.sa_out_w, L.sa_out_b, 0, nullptr, proj, Rq, D, D, ACT_NONE));
      MOCHA_TRY(add_layernorm(dx, proj, L.n1_g, L.n1_b, dy, Rq, D, w->ln_eps, nullptr, nullptr, 0, nullptr, c.s));
    }
    // cross-attention over the memory
    MOCHA_TRY(dense(c, dy, D, L.ca_in_w, L.ca_in_b, 0, nullptr, dq, Rq, D, D, ACT_NONE));
    MOCHA_TRY(dense(c, mem, D, L.ca_in_w + (size_t)D * D, L.ca_in_b + D, 0, nullptr, memkv, Rm, 2 * D, D, ACT_NONE));
    MOCHA_TRY(attention(c, dq, D, memkv, 2 * D, memkv + D, 2 * D, B, H, nq, nm, dh, S, att, D));
    MOCHA_TRY(dense(c, att, D, L.ca_out_w, L.ca_out_b, 0, nullptr, proj, Rq, D, D, ACT_NONE));
    MOCHA_TRY(add_layernorm(dy, proj, L.n2_g, L.n2_b, dx, Rq, D, w->ln_eps, nullptr, nullptr, 0, nullptr, c.s));
    // feed-forward
    MOCHA_TRY(dense(c, dx, D, L.l1_w, L.l1_b, 0, nullptr, hid, Rq, w->dff, D, ACT_RELU));
    MOCHA_TRY(dense(c, hid, w->dff, L.l2_w, L.l2_b, 0, nullptr, proj, Rq, D, w->dff, ACT_NONE));
    if (l == w->depth - 1) {
      MOCHA_TRY(add_layernorm(dx, proj, L.n3_g, L.n3_b, out, Rq, D, w->ln_eps, out_mean, out_std, nq, out_denorm, c.s));
    } else {
      MOCHA_TRY(add_layernorm(dx, proj, L.n3_g, L.n3_b, dy, Rq, D, w->ln_eps, nullptr, nullptr, 0, nullptr, c.s));
      float* t = dx; dx = dy; dy = t;
    }
  }
  return MOCHA_OK;
}

extern "C" int mocha_cvae_precompute_dec0(const mocha_cvae_weights* w, float* table, void* workspace,
                                          size_t workspace_bytes, mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && table, "mocha_cvae_precompute_dec0: null argument");
  MOCHA_CHECK_ARG(w->D > 0 && w->heads > 0 && w->D % w->heads == 0 && w->out_seq > 0 && w->depth >= 1,
                  "mocha_cvae_precompute_dec0: bad geometry");
  Workspace ws(workspace, workspace_bytes);
  Ctx c{(cudaStream_t)stream, MOCHA_FP32, &ws};
  const int D = w->D, H = w->heads, dh = D / H, nq = w->out_seq;
  float* qkv = ws.take<float>((size_t)nq * 3 * D);
  float* S = ws.take<float>((size_t)H * nq * nq);
  float* att = ws.take<float>((size_t)nq * D);
  float* proj = ws.take<float>((size_t)nq * D);
  WS_GUARD(ws, "mocha_cvae_precompute_dec0");
  const mocha_cvae_dec_layer& L = w->dec[0];
  MOCHA_CHECK_ARG(L.sa_in_w && L.sa_in_b && L.sa_out_w && L.sa_out_b && L.n1_g && L.n1_b, "mocha_cvae_precompute_dec0: weights missing");
  MOCHA_TRY(dense(c, w->pe, D, L.sa_in_w, L.sa_in_b, 0, nullptr, qkv, nq, 3 * D, D, ACT_NONE));
  MOCHA_TRY(attention(c, qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, 1, H, nq, nq, dh, S, att, D));
  MOCHA_TRY(dense(c, att, D, L.sa_out_w, L.sa_out_b, 0, nullptr, proj, nq, D, D, ACT_NONE));
  return add_layernorm(w->pe, proj, L.n1_g, L.n1_b, table, nq, D, w->ln_eps, nullptr, nullptr, 0, nullptr, c.s);
}

extern "C" int mocha_cvae_condition(const float* src_cnt, const float* prev, const float* m0, const float* s0,
                                    const float* m1, const float* s1, float* cond, int B, int n, int D,
                                    mocha_stream_t stream) {
  return cvae_condition(src_cnt, prev, m0, s0, m1, s1, cond, B, n, D, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// exposed dense primitive
// ------------------------------------------------------------------------------------------------
extern "C" size_t mocha_linear_workspace_bytes(int M, int N, int K, int precision) {
  if (precision == MOCHA_TF32X3) return ((size_t)M + (size_t)N) * 3 * (size_t)K * 4 + 8192;
  return precision == MOCHA_BF16 ? tc_scratch_bytes((size_t)M, (size_t)K) + 4096 : 0;
}

extern "C" int mocha_linear(const float* A, const float* W, const float* bias, const float* res, float* C, int M,
                            int N, int K, int act, int precision, void* workspace, size_t workspace_bytes,
                            mocha_stream_t stream) {
  MOCHA_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0, "mocha_linear: null/empty argument");
  MOCHA_CHECK_ARG(act >= 0 && act <= 3, "mocha_linear: unknown activation %d", act);
  Workspace ws(workspace, workspace_bytes);
  Ctx c{(cudaStream_t)stream, precision, &ws};
  int rc = dense(c, A, K, W, bias, 0, res, C, M, N, K, act);
  if (rc == MOCHA_OK && ws.overflow)
    return set_error(MOCHA_ERR_WORKSPACE, "mocha_linear: workspace too small");
  return rc;
}

// Stand-alone entry for the dominant kernel of the batched path (bench.py roofline pass): the
// reflect-padded temporal convolution of mot_embedding's JointBlock, x [B*T*V, D] -> out [B*T*V, D].
extern "C" int mocha_bench_tconv(const mocha_generator_weights* w, const float* x, int B, float* out, int precision,
                                 int gemm_repeats, void* workspace, size_t workspace_bytes, mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && x && out && B > 0, "mocha_bench_tconv: null/empty argument");
  MOCHA_TRY(check_dims(w->dims));
  const mocha_dims& d = w->dims;
  Workspace ws(workspace, workspace_bytes);
  Ctx c{(cudaStream_t)stream, precision, &ws};
  if (precision == MOCHA_BF16 && tc_tconv_supported(B, d.T, d.V, d.D, d.D, d.taps_j))
    return tc_tconv(x, w->jb_tcn_w, w->jb_tcn_b, 0, out, B, d.T, d.V, d.D, d.D, d.taps_j, 1, ws, c.s, gemm_repeats);
  int rc = MOCHA_OK;
  for (int it = 0; it < (gemm_repeats < 1 ? 1 : gemm_repeats) && rc == MOCHA_OK; ++it)
    rc = tconv(c, x, w->jb_tcn_w, w->jb_tcn_b, 0, out, B, d.T, d.V, d.D, d.D, d.taps_j, 1);
  return rc;
}

// Stand-alone launches of the bandwidth-bound kernels of the batched path at the step's shapes (bench.py
// `hbm_kernels`): `repeats` back-to-back launches on the caller's stream; *algo_bytes receives the ALGORITHMIC bytes
// of one launch (every input element read once + every output element written once).
// which: 0 embed_graph_agg (1x1 embed conv + LeakyReLU + joint-graph aggregation), 1 pool_graph_agg,
// 2 add_layernorm (CVAE prior rows, fp32 + bf16 outputs), 3 graph_agg_kv_pad16 (to_mot), 4 adain_norm_tokens,
// 5 instance_norm_tokens -> bf16 (decoder style tokens), 6 out_conv_affine (to_mot output layer + de-normalisation),
// 7 graph_agg_small (to_mot body-part aggregation)
extern "C" int mocha_bench_hbm_kernel(const mocha_generator_weights* w, int which, int B, int repeats, void* workspace,
                                      size_t workspace_bytes, double* algo_bytes, mocha_stream_t stream) {
  MOCHA_CHECK_ARG(w && B > 0 && repeats >= 1 && algo_bytes, "mocha_bench_hbm_kernel: bad argument");
  MOCHA_TRY(check_dims(w->dims));
  const mocha_dims& d = w->dims;
  Workspace ws(workspace, workspace_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  const int R = B * d.T * d.V, Tp = d.T / d.tp, R2 = B * Tp * d.P, n = Tp * d.P;
  typedef __nv_bfloat16 bf16;
  int rc = MOCHA_OK;
  switch (which) {
    case 0: {
      const int KC = d.Kj * d.C0;
      const int Ka = (w->jb_gcn_w_aug && w->jb_gcn_kaug >= KC + d.Kj) ? w->jb_gcn_kaug : KC;
      float* X = ws.take<float>((size_t)R * d.Cin);
      bf16* out = ws.take<bf16>((size_t)R * Ka);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      MOCHA_CUDA(cudaMemsetAsync(X, 0, (size_t)R * d.Cin * 4, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = embed_graph_agg(X, w->emb_w, w->emb_b, w->A_j, out, B * d.T, d.V, d.Cin, d.C0, d.Kj, s, Ka == KC ? 0 : Ka);
      *algo_bytes = (double)R * d.Cin * 4 + (double)R * Ka * 2;
      break;
    }
    case 1: {
      bf16* h1 = ws.take<bf16>((size_t)R * d.D);
      bf16* out = ws.take<bf16>((size_t)R2 * d.Kb * d.D);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      MOCHA_CUDA(cudaMemsetAsync(h1, 0, (size_t)R * d.D * 2, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = pool_graph_agg(h1, w->pool_w, w->A_b, out, B, d.T, d.V, d.P, d.D, d.tp, d.Kb, s);
      *algo_bytes = (double)R * d.D * 2 + (double)R2 * d.Kb * d.D * 2;
      break;
    }
    case 2: {
      const long long rows = (long long)B * (2 * n + 2);
      float* x = ws.take<float>((size_t)rows * d.D);
      float* y = ws.take<float>((size_t)rows * d.D);
      bf16* y16 = ws.take<bf16>((size_t)rows * d.D);
      float* g = ws.take<float>((size_t)2 * d.D);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      MOCHA_CUDA(cudaMemsetAsync(x, 0, (size_t)rows * d.D * 4, s));
      MOCHA_CUDA(cudaMemsetAsync(g, 0, (size_t)2 * d.D * 4, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = add_layernorm(x, nullptr, g, g + d.D, y, rows, d.D, 1e-5f, nullptr, nullptr, 0, nullptr, s, y16);
      *algo_bytes = (double)rows * d.D * (4 + 4 + 2);
      break;
    }
    case 3: {
      const int padj = d.taps_j / 2;
      float* y3 = ws.take<float>((size_t)R2 * d.Kj * d.C0);
      bf16* gp = ws.take<bf16>((size_t)B * (d.T + 2 * padj) * d.V * d.C0);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      MOCHA_CUDA(cudaMemsetAsync(y3, 0, (size_t)R2 * d.Kj * d.C0 * 4, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = graph_agg_kv_pad16(y3, w->tm_A2, gp, B, Tp, d.tp, padj, d.P, d.V, d.C0, d.Kj, s);
      *algo_bytes = (double)R2 * d.Kj * d.C0 * 4 + (double)B * (d.T + 2 * padj) * d.V * d.C0 * 2;
      break;
    }
    case 4: case 5: {
      float* x = ws.take<float>((size_t)B * n * d.D);
      float* gb = ws.take<float>((size_t)B * 2 * d.D);
      float* y = ws.take<float>((size_t)B * n * d.D);
      bf16* q16 = ws.take<bf16>((size_t)B * n * d.D);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      // a ramp, not zeros: a constant channel has zero variance
      MOCHA_TRY(broadcast_rows(w->pos_emb, x, B, (long long)n * d.D, s));
      MOCHA_CUDA(cudaMemsetAsync(gb, 0, (size_t)B * 2 * d.D * 4, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = which == 4 ? adain_norm_tokens(x, B, n, d.D, 1e-5f, gb, y, q16, s)
                        : instance_norm_tokens(x, B, n, d.D, 1e-5f, nullptr, nullptr, nullptr, nullptr, nullptr, s, q16);
      *algo_bytes = which == 4 ? (double)B * n * d.D * (4 + 4 + 2) : (double)B * n * d.D * (4 + 2);
      break;
    }
    case 6: {
      // to_mot output layer + de-normalisation (out_conv_affine): bf16 rows in, de-normalised fp32 Y out
      bf16* y4 = ws.take<bf16>((size_t)R * d.C0);
      float* Y = ws.take<float>((size_t)R * d.Cin);
      float* tab = ws.take<float>((size_t)2 * d.V * d.Cin);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      MOCHA_CHECK_ARG(out_conv_affine_supported(d.C0, d.Cin), "mocha_bench_hbm_kernel: out_conv_affine does not take this geometry");
      MOCHA_CUDA(cudaMemsetAsync(y4, 0, (size_t)R * d.C0 * 2, s));
      MOCHA_CUDA(cudaMemsetAsync(tab, 0, (size_t)2 * d.V * d.Cin * 4, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = out_conv_affine(y4, w->tm_out_w, w->tm_out_b, tab, tab + d.V * d.Cin, nullptr, Y, R, d.C0, d.Cin, d.V, s);
      *algo_bytes = (double)R * d.C0 * 2 + (double)R * d.Cin * 4;
      break;
    }
    case 7: {
      // to_mot's body-part graph aggregation (graph_agg_small): fp32 tokens in, bf16 [R2, Kb*D] out
      float* x = ws.take<float>((size_t)R2 * d.D);
      bf16* out = ws.take<bf16>((size_t)R2 * d.Kb * d.D);
      WS_GUARD(ws, "mocha_bench_hbm_kernel");
      MOCHA_CUDA(cudaMemsetAsync(x, 0, (size_t)R2 * d.D * 4, s));
      for (int i = 0; i < repeats && rc == MOCHA_OK; ++i)
        rc = graph_agg_first(x, w->tm_A_b, nullptr, B * Tp, d.P, d.D, d.Kb, 1, s, out);
      *algo_bytes = (double)R2 * d.D * 4 + (double)R2 * d.Kb * d.D * 2;
      break;
    }
    default:
      return set_error(MOCHA_ERR_ARG, "mocha_bench_hbm_kernel: unknown kernel %d", which);
  }
  return rc;
}

// Stand-alone entry of the fused block-tail kernel (unit tests, bench.py): operands already bf16.
extern "C" int mocha_block_tail(const void* A0, int lda, int K0, const void* W0, const float* b0, const float* R0,
                                const float* g1, const float* be1, int Hd, int act, const void* W1, const float* b1,
                                const void* W2, const float* b2, const float* g2, const float* be2, float eps, float* O32,
                                void* O16, int M, mocha_stream_t stream) {
  return tc_tail((const __nv_bfloat16*)A0, lda, K0, (const __nv_bfloat16*)W0, b0, R0, g1, be1, Hd, act,
                 (const __nv_bfloat16*)W1, b1, (const __nv_bfloat16*)W2, b2, g2, be2, eps, O32, (__nv_bfloat16*)O16, M,
                 (cudaStream_t)stream);
}

// Stand-alone entry of the fused attention core (unit tests, bench.py): bf16 q / k / v views, bf16 output.
extern "C" int mocha_attention_core(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int B, int H, int nq,
                                    int nkv, int dh, void* out, int ldo, mocha_stream_t stream) {
  return tc_attn_fused((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, B, H, nq, nkv, dh,
                       (__nv_bfloat16*)out, ldo, (cudaStream_t)stream);
}
