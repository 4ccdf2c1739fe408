/* TEST INFRASTRUCTURE — declaration-only stand-in for Intel MKL's header, so that the reference's
 * DiskANN/src/pq.cpp compiles here.  Only the PQ *query* functions of that file are ever called by the
 * harness (load_pq_centroid_bin, preprocess_query, populate_chunk_distances, aggregate_coords,
 * pq_dist_lookup); they use no BLAS/LAPACK.  The training functions that do are left unresolved on purpose
 * and abort if reached (diskann_ref_harness.cpp). */
#pragma once
#include <cstddef>
typedef int MKL_INT;
enum CBLAS_LAYOUT { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112 };
#define LAPACK_ROW_MAJOR 101
extern "C" {
void cblas_sgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, MKL_INT, MKL_INT, MKL_INT, float, const float*, MKL_INT,
                 const float*, MKL_INT, float, float*, MKL_INT);
float cblas_snrm2(MKL_INT, const float*, MKL_INT);
float cblas_sdot(MKL_INT, const float*, MKL_INT, const float*, MKL_INT);
int LAPACKE_sgesdd(int, char, MKL_INT, MKL_INT, float*, MKL_INT, float*, float*, MKL_INT, float*, MKL_INT);
}
