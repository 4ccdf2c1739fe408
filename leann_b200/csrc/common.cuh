// Internal declarations shared by the translation units of libleann_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb2 {

// thread-local last-error string behind lb2_last_error()
void set_error(const char* fmt, ...);

#define LB2_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            lb2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return false;                                                              \
        }                                                                              \
    } while (0)

enum Epilogue { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RES = 2 };

// Epilogue operands of the tcgen05 GEMM.  With the "deferred LayerNorm" fields set, the LayerNorms of
// the BERT block never run as kernels: a producer GEMM (EPI_BIAS_RES) emits per-row partial
// (sum, sum of squares) of the pre-LN tensor it writes, and the consumers apply the normalisation
// algebraically —
//   LN-in  (EPI_BIAS / EPI_BIAS_GELU, stats_in != null):  the GEMM ran on the PRE-LN rows y with weights
//           W' = W * gamma folded in, so   out = r * (acc - mu * svec[n]) + bias[n],
//           svec[n] = sum_k W'[n,k],  bias[n] = sum_k beta[k] W[n,k] + b[n]   (precomputed at load time);
//   LN-res (EPI_BIAS_RES, ln_g != null):  the residual tile holds pre-LN rows p, used as
//           (p - mu) * r * gamma[n] + beta[n].
// Partials are kept in fixed slots and summed in a fixed order (no atomics): results stay bit-reproducible.
struct EpiArgs {
    const float* bias = nullptr;      // [N]
    const float* svec = nullptr;      // [N]   LN-in only
    const float* stats_in = nullptr;  // [M][2 * parts_in]  stats of the A rows (LN-in) or of the residual rows (LN-res)
    int parts_in = 0;
    const float* ln_g = nullptr;      // [N]   LN-res only
    const float* ln_b = nullptr;      // [N]
    float* stats_out = nullptr;       // [M][2 * (N / 96)]  EPI_BIAS_RES only: partial stats of the rows written
    float inv_width = 0.f;            // 1 / (normalised width of the incoming LN)
    float eps = 1e-12f;
};

// ---- gemm_tcgen05.cu
bool make_tmap_f16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols = 64);
bool gemm_f16(cudaStream_t stream, const __half* A, const CUtensorMap* tmap_w, const __half* W, const EpiArgs& ea,
              const __half* residual, __half* C, int M, int N, int K, int epi, int num_sms, int c_group = 0);
int gemm_block_n();
// optional per-kernel-class timing with CUDA events (api.cu owns the pools; no-ops unless
// LB2_PROFILE_GEMM=1): category 0 = tcgen05 GEMMs, 1 = attention, 2 = LayerNorm / embed / pool
enum ProfCat { PROF_GEMM = 0, PROF_ATTN = 1, PROF_NORM = 2, PROF_NCAT = 3 };
void prof_begin(cudaStream_t s, int cat);
void prof_end(cudaStream_t s, int cat, double flops);

// ---- encoder.cu
struct EncoderConfig {
    int vocab_size, hidden, layers, heads, ffn, max_pos, type_vocab;
    float ln_eps;
    int pooling;    // 0 = masked mean (all-MiniLM-L6-v2), 1 = CLS (bge-base-en-v1.5)
    int normalize;  // 1 = L2-normalise the pooled vector (both models ship a Normalize module)
};

struct LayerWeights {
    __half* w_qkv;  // [3H, H]  rows regrouped per head (q_h | k_h | v_h), gamma of the incoming LN folded in
    float* t_qkv;   // [3H]     folded bias: sum_k beta[k] W[n,k] + b[n]
    float* s_qkv;   // [3H]     row sums of the folded (fp16) weights
    __half* w_o;    // [H, H]
    float* b_o;
    float *ln1_g, *ln1_b;
    __half* w_1;  // [F, H]   gamma of LN1 folded in
    float *t_1, *s_1;
    __half* w_2;  // [H, F]
    float* b_2;
    float *ln2_g, *ln2_b;
    CUtensorMap tm_qkv, tm_o, tm_1, tm_2;
};

struct Encoder {
    EncoderConfig cfg{};
    bool loaded = false;
    __half *word_emb = nullptr, *pos_emb = nullptr, *type_emb = nullptr;
    float *emb_ln_g = nullptr, *emb_ln_b = nullptr;
    LayerWeights* layers = nullptr;  // host array of device pointers
    void* arena = nullptr;           // one device allocation holding every tensor above
    // workspaces, sized for `cap_tokens` packed tokens / `cap_seqs` passages.  p[] are the PRE-LayerNorm
    // residual-stream tensors (the LayerNorms are applied algebraically by their consumers, see EpiArgs),
    // st[] the per-row partial (sum, sum of squares) that go with them.
    int64_t cap_tokens = 0, cap_seqs = 0;
    __half* p[3] = {nullptr, nullptr, nullptr};
    float* st[3] = {nullptr, nullptr, nullptr};
    __half *qkv = nullptr, *ctx = nullptr, *ffn = nullptr;
    int32_t* seq_len = nullptr;  // [cap_seqs] truncated passage lengths of the current pass
    int parts = 4;               // partial-statistics slots per row = hidden / 96
    int num_sms = 148;
};

size_t encoder_weight_floats(const EncoderConfig& cfg);
bool encoder_load(Encoder* enc, const EncoderConfig& cfg, const float* host_weights, size_t n_floats);
void encoder_free(Encoder* enc);
bool encoder_reserve(Encoder* enc, int64_t tokens, int64_t seqs);
// Encode n_seq passages.  Passage i is node seq_node[i] of the token store (tok_store / tok_off,
// device pointers); its rows in the packed activation matrix start at seq_start[i] - row_base;
// n_tokens = packed row count.  Writes pooled (+normalised) fp32 embeddings to out[n_seq, H].
bool encoder_forward(Encoder* enc, cudaStream_t stream, const uint16_t* tok_store, const uint64_t* tok_off,
                     const int32_t* seq_node, const int32_t* seq_start, int row_base, int n_seq, int n_tokens,
                     float* out);

// kernels exposed for the unit-test hooks in api.cu
// qkv is head-major: [heads][n_tokens][3 * head_dim] (q | k | v per token)
bool launch_attention(cudaStream_t s, const __half* qkv, const int32_t* seq_start, const int32_t* seq_len, int row_base,
                      int max_pos, int n_seq, int n_tokens, int hidden, int heads, __half* ctx);

}  // namespace lb2
