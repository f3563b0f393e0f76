// TEST HARNESS (host build of csrc/strong_math.cuh): the element arithmetic of the strongly periodic equation,
// exactly the functions the CUDA kernels of csrc/strong.cu call, exposed to tests/test_oracle_strong.py so that it
// can be compared with the oracle on a box without a GPU.  Dense outputs, one block per cell / per (cell, facet).
#include "../dmri-fem-cloud_b200/csrc/strong_math.cuh"

extern "C" {

// W[t][i][j], C[t][i][j] for all cells (4x4 blocks, unused slots zero)
void strong_host_cells(int64_t nc, int cell_nv, const double* xyz, const int32_t* cells4, int dkind, const double* D,
                       const double* g, double* W, double* C) {
  for (int64_t t = 0; t < nc; ++t)
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double w = 0.0, c = 0.0;
        if (i < cell_nv && j < cell_nv) strong_cell_wc(xyz, cells4, t, cell_nv, i, j, dkind, D, g, &w, &c);
        W[(t * 4 + i) * 4 + j] = w;
        C[(t * 4 + i) * 4 + j] = c;
      }
}

// coef[t][lf] of every (cell, local facet)
void strong_host_facets(int64_t nc, int cell_nv, const double* xyz, const int32_t* cells4, int dkind, const double* D,
                        const double* g, double* coef) {
  for (int64_t t = 0; t < nc; ++t)
    for (int lf = 0; lf < 4; ++lf)
      coef[t * 4 + lf] = lf < cell_nv ? strong_facet_coef(xyz, cells4, t, cell_nv, lf, dkind, D, g) : 0.0;
}

void strong_host_combine(double mk, double k0t, double w, double gv, double aA, double aP, double di, double* out3) {
  strong_combine_entry(mk, k0t, w, gv, aA, aP, di, &out3[0], &out3[1], &out3[2]);
}
}
