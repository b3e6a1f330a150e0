// MOCHA_BF16 stage bodies (see networks_bf16.cu). Called by the C-ABI stage functions in networks.cu.
#pragma once
#include "common.cuh"

namespace mocha {

bool bf16_path_supported(const mocha_dims& d);
int embed_bf16(const mocha_generator_weights* w, const float* X, int B, float* tokens, int add_pos_emb, Workspace& ws,
               cudaStream_t s);
int encoder_bf16(const mocha_generator_weights* w, const float* tokens, int B, float* encoded, Workspace& ws,
                 cudaStream_t s);
int decoder_bf16(const mocha_generator_weights* w, const float* src, const float* cha, int B, float* decoded,
                 Workspace& ws, cudaStream_t s);
int to_mot_bf16(const mocha_generator_weights* w, const float* tokens, int B, float* Ytil, const float* Y_mean,
                const float* Y_std, float* Y, Workspace& ws, cudaStream_t s);
int cvae_bf16(const mocha_cvae_weights* w, const float* cond, int B, int ncond, const float* eps, float* out, float* mu,
              float* logvar, const float* out_mean, const float* out_std, float* out_denorm, Workspace& ws,
              cudaStream_t s);

}  // namespace mocha
