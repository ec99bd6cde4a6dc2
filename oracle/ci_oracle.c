/* Oracle (TEST INFRASTRUCTURE): integer-exact CPU restatement of the cluster-index
 * sphere-growing search of the reference, CI.py:87-145 (calculate_CV / calculate_CI)
 * with px2vec CI.py:65-68 and getRadiiIndices CI.py:79-85.
 *
 * PINNED: tests/test_oracle_golden.py checks it bit-for-bit against outputs of the
 * reference's own CI.calculate_CI (tests/golden/ci_*.npz) and against the literal
 * numpy restatement in oracle/ci_oracle.py.
 *
 * Restatement.  px2vec maps (i,j,k) -> i + (j-1)*s0 + (k-1)*s0*s1 and is applied to
 * UNCHECKED neighbour coordinates (CI.py:93), so the sphere lives in flat,
 * row-fastest index space: f = i + j*s0 + k*s0*s1, neighbour f' = f + dx + dy*s0 +
 * dz*s0*s1, a hit iff 0 <= f' < s0*s1*s2 and flat voxel f' is a defect.  For the
 * shell boundaries ii (row index where the radius increases), CI.py:96-99 stops at the
 * first ii with len(intersect1d(sphereVec[:ii], defVec))/ii < 0.5, i.e.
 * 2*|unique hits among rows < ii| < ii; the value is table[ii-1][0] (CI.py:105).
 * No boundary stops -> ValueError in the reference (CI.py:101-103) -> rc 1 here.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* defect: [n0][n1][n2] C order (the numpy array), non-zero = defect.
 * table:  nrows x 4 doubles [radius, dx, dy, dz].
 * shell_out: [n0][n1][n2] int32, boundary ordinal t (0-based index into the list of
 *            boundaries) at which the voxel stopped; -1 off-defect.
 * cv_out:    [n0][n1][n2] double, table[ii_t - 1][0] (NOT yet multiplied by min(vox)).
 * returns 0, or 1 if some voxel exhausted the table (reference raises ValueError). */
int ci_oracle(const uint8_t *defect, int n0, int n1, int n2,
              const double *table, int64_t nrows,
              int32_t *shell_out, double *cv_out)
{
    const int64_t N = (int64_t)n0 * n1 * n2;
    uint8_t *flat = (uint8_t *)calloc((size_t)N + 1, 1);
    int32_t *stamp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)nrows);
    int64_t *bnd = (int64_t *)malloc(sizeof(int64_t) * (size_t)nrows);
    int64_t nb = 0;
    int rc = 0;
    for (int64_t i = 0; i <= N; ++i) stamp[i] = -1;
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j)
            for (int k = 0; k < n2; ++k)
                if (defect[((int64_t)i * n1 + j) * n2 + k])
                    flat[i + (int64_t)j * n0 + (int64_t)k * n0 * n1] = 1;
    for (int64_t r = 0; r < nrows; ++r)
        off[r] = (int64_t)table[4 * r + 1] + (int64_t)table[4 * r + 2] * n0 +
                 (int64_t)table[4 * r + 3] * n0 * n1;
    for (int64_t r = 0; r + 1 < nrows; ++r)           /* getRadiiIndices: where(diff>0)+1 */
        if (table[4 * (r + 1)] - table[4 * r] > 0) bnd[nb++] = r + 1;

    int32_t id = 0;
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j)
            for (int k = 0; k < n2; ++k) {
                const int64_t c = ((int64_t)i * n1 + j) * n2 + k;
                shell_out[c] = -1;
                cv_out[c] = 0.0;
                if (!defect[c]) continue;
                const int64_t f = i + (int64_t)j * n0 + (int64_t)k * n0 * n1;
                int64_t row = 0, num = 0, t;
                for (t = 0; t < nb; ++t) {
                    const int64_t ii = bnd[t];
                    for (; row < ii; ++row) {
                        const int64_t g = f + off[row];
                        if (g >= 0 && g < N && flat[g] && stamp[g] != id) {
                            stamp[g] = id;            /* intersect1d counts unique voxels */
                            ++num;
                        }
                    }
                    if (2 * num < ii) break;          /* C < 0.5 */
                }
                if (t == nb) { rc = 1; ++id; continue; }
                shell_out[c] = (int32_t)t;
                cv_out[c] = table[4 * (bnd[t] - 1)];
                ++id;
            }
    free(flat); free(stamp); free(off); free(bnd);
    return rc;
}
