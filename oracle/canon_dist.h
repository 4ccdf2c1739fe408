/* TEST INFRASTRUCTURE (oracle/) — never linked into the product library.
 *
 * Canonical fp32 distance arithmetic shared by the C oracle and the harness
 * around the compiled reference.
 *
 * The reference scores a hop's batch with  -E@q  (mips/cosine) or
 * sum((E-q)^2) (l2) in numpy
 * (packages/leann-backend-hnsw/leann_backend_hnsw/hnsw_embedding_server.py:195-200)
 * and, for single ids, with faiss' scalar loops fvec_inner_product / fvec_L2sqr
 * (third_party/faiss/faiss/utils/distances_simd.cpp:194-203, 219-230).  Both
 * leave the fp32 summation ORDER unspecified (BLAS / "imprecise loop" pragma),
 * so bit-exact traversal parity needs one agreed order.  We fix it to the
 * order the CUDA warp uses:
 *   lane l (0..31) accumulates elements l, l+32, l+64, ... with fused
 *   multiply-add in increasing index order, then the 32 partials are combined
 *   by an xor-butterfly (offsets 16, 8, 4, 2, 1).
 * fp32 '+' is commutative, so every lane ends with the same value.
 */
#ifndef LB2_CANON_DIST_H
#define LB2_CANON_DIST_H
#include <math.h>

static inline float lb2_canon_reduce32(float p[32]) {
    for (int off = 16; off >= 1; off >>= 1) {
        float t[32];
        for (int l = 0; l < 32; l++) t[l] = p[l] + p[l ^ off];
        for (int l = 0; l < 32; l++) p[l] = t[l];
    }
    return p[0];
}

static inline float lb2_canon_ip(const float* a, const float* b, int d) {
    float p[32];
    for (int l = 0; l < 32; l++) {
        float acc = 0.0f;
        for (int j = l; j < d; j += 32) acc = fmaf(a[j], b[j], acc);
        p[l] = acc;
    }
    return lb2_canon_reduce32(p);
}

static inline float lb2_canon_l2(const float* a, const float* b, int d) {
    float p[32];
    for (int l = 0; l < 32; l++) {
        float acc = 0.0f;
        for (int j = l; j < d; j += 32) {
            float t = a[j] - b[j];
            acc = fmaf(t, t, acc);
        }
        p[l] = acc;
    }
    return lb2_canon_reduce32(p);
}
#endif
