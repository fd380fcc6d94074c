// qnb_shake.cuh -- SHAKE for solvent-sized molecules (the per-step host neighbour of the nonbonded call).
//
// shake(xx, x) of bondene.f90:1069-1150 over the constraints init_constraints (simprep.f90:2167-2345) builds: every
// constrained molecule is relaxed on its own -- sweep its constraints in list order, correct x(i), x(j) along the
// REFERENCE bond vector xx(i)-xx(j), flag a constraint ready the first time it is found within CONST_TOL (it still
// receives that sweep's correction and is not looked at again), stop when all are flagged.  Molecules do not share
// atoms, so one THREAD per molecule runs the reference's sequential algorithm literally (same order of operations,
// explicitly rounded FP64: the result is bit-identical to the CPU restatement), reading and writing only its own atoms.
// A water (3 constraints) needs 2-10 sweeps; the kernel is latency-bound and tiny next to the force kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace qnb {

constexpr double kConstTol = 0.0001;   // CONST_TOL, globals.f90:519
constexpr int kConstMaxIter = 1000;    // CONST_MAX_ITER, globals.f90:520
constexpr int kMaxMolConstraints = 32; // ready flags of one molecule live in a 32-bit mask

__global__ void __launch_bounds__(128)
k_shake(int nmol, const int *__restrict__ mol_first, const int2 *__restrict__ cij, const double *__restrict__ dist2,
        const double *__restrict__ winv, const double *__restrict__ xx, double *__restrict__ x,
        unsigned long long *__restrict__ iter_sum, int *__restrict__ failed) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    const int c0 = mol_first[m], c1 = mol_first[m + 1];
    const int nc = c1 - c0;
    if (nc <= 0) return;
    const unsigned all = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
    unsigned ready = 0;
    int nits = 0;
    for (;;) {
        for (int c = c0; c < c1; c++) {
            const unsigned bit = 1u << (c - c0);
            if (ready & bit) continue;
            const int2 a = cij[c];
            const int i = 3 * a.x, j = 3 * a.y;
            const double d2 = dist2[c];
            const double xi0 = x[i], xi1 = x[i + 1], xi2 = x[i + 2];
            const double xj0 = x[j], xj1 = x[j + 1], xj2 = x[j + 2];
            // xij = q_dist5(x(j), x(i))%vec = x(i) - x(j)   (math.f90:270)
            const double v0 = __dsub_rn(xi0, xj0), v1 = __dsub_rn(xi1, xj1), v2 = __dsub_rn(xi2, xj2);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1)), __dmul_rn(v2, v2));
            const double diff = __dsub_rn(d2, r2);
            if (fabs(diff) < __dmul_rn(kConstTol, d2)) ready |= bit;
            const double w0 = __dsub_rn(xx[i], xx[j]), w1 = __dsub_rn(xx[i + 1], xx[j + 1]), w2 = __dsub_rn(xx[i + 2], xx[j + 2]);
            const double scp = __dadd_rn(__dadd_rn(__dmul_rn(v0, w0), __dmul_rn(v1, w1)), __dmul_rn(v2, w2));
            const double wi = winv[a.x], wj = winv[a.y];
            const double corr = __ddiv_rn(diff, __dmul_rn(__dmul_rn(2.0, scp), __dadd_rn(wi, wj)));
            // x(i) = x(i) + xxij*corr*winv(i);  x(j) = x(j) + (-xxij)*corr*winv(j)
            x[i] = __dadd_rn(xi0, __dmul_rn(__dmul_rn(w0, corr), wi));
            x[i + 1] = __dadd_rn(xi1, __dmul_rn(__dmul_rn(w1, corr), wi));
            x[i + 2] = __dadd_rn(xi2, __dmul_rn(__dmul_rn(w2, corr), wi));
            x[j] = __dadd_rn(xj0, __dmul_rn(__dmul_rn(-w0, corr), wj));
            x[j + 1] = __dadd_rn(xj1, __dmul_rn(__dmul_rn(-w1, corr), wj));
            x[j + 2] = __dadd_rn(xj2, __dmul_rn(__dmul_rn(-w2, corr), wj));
        }
        nits++;
        if (ready == all) break;
        if (nits >= kConstMaxIter) {   // die('shake failure'), bondene.f90:1143
            atomicExch(failed, 1);
            break;
        }
    }
    atomicAdd(iter_sum, (unsigned long long)nits);
}

}  // namespace qnb
