// qnb_shake.cuh -- SHAKE for solvent-sized molecules (the per-step host neighbour of the nonbonded call).
//
// shake(xx, x) of bondene.f90:1069-1150 over the constraints init_constraints (simprep.f90:2167-2345) builds: every
// constrained molecule is relaxed on its own -- sweep its constraints in list order, correct x(i), x(j) along the
// REFERENCE bond vector xx(i)-xx(j), flag a constraint ready the first time it is found within CONST_TOL (it still
// receives that sweep's correction and is not looked at again), stop when all are flagged.  Molecules do not share
// atoms, so one THREAD per molecule runs the reference's sequential algorithm literally (same order of operations,
// explicitly rounded FP64: the result is bit-identical to the CPU restatement), reading and writing only its own atoms.
// A water (3 constraints) needs 2-10 sweeps; the kernel is latency-bound and tiny next to the force kernels.
//
// shake_molecule() is __host__ __device__: tests/cpu_shims/shake_shim.cpp compiles THIS source with g++ and the CPU suite
// runs it against the restatement (tests/test_shake_cpu.py), so the algorithm the kernel executes is checked without a
// GPU; only the thread mapping and the ABI plumbing in qnb.cu need hardware.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define QNB_SHAKE_HD __host__ __device__ __forceinline__
#else
#define QNB_SHAKE_HD inline
#endif

namespace qnb {

constexpr double kConstTol = 0.0001;   // CONST_TOL, globals.f90:519
constexpr int kConstMaxIter = 1000;    // CONST_MAX_ITER, globals.f90:520
constexpr int kMaxMolConstraints = 32; // ready flags of one molecule live in a 32-bit mask

struct ShakePair { int i, j; };        // 0-based atoms of one constraint

// Individually rounded FP64 operations: intrinsics on the device (no FMA contraction whatever the compiler flags),
// plain operators on the host (the shim is built with -ffp-contract=off).
QNB_SHAKE_HD double sh_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
QNB_SHAKE_HD double sh_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
QNB_SHAKE_HD double sh_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
QNB_SHAKE_HD double sh_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// One molecule: constraints [c0, c1).  Returns the number of sweeps; *failed is set when CONST_MAX_ITER is reached.
QNB_SHAKE_HD int shake_molecule(int c0, int c1, const ShakePair *cij, const double *dist2, const double *winv,
                                const double *xx, double *x, bool *failed) {
    const int nc = c1 - c0;
    if (nc <= 0) return 0;
    const unsigned all = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
    unsigned ready = 0;
    int nits = 0;
    for (;;) {
        for (int c = c0; c < c1; c++) {
            const unsigned bit = 1u << (c - c0);
            if (ready & bit) continue;
            const ShakePair a = cij[c];
            const int i = 3 * a.i, j = 3 * a.j;
            const double d2 = dist2[c];
            const double xi0 = x[i], xi1 = x[i + 1], xi2 = x[i + 2];
            const double xj0 = x[j], xj1 = x[j + 1], xj2 = x[j + 2];
            // xij = q_dist5(x(j), x(i))%vec = x(i) - x(j)   (math.f90:270)
            const double v0 = sh_sub(xi0, xj0), v1 = sh_sub(xi1, xj1), v2 = sh_sub(xi2, xj2);
            const double r2 = sh_add(sh_add(sh_mul(v0, v0), sh_mul(v1, v1)), sh_mul(v2, v2));
            const double diff = sh_sub(d2, r2);
            if (fabs(diff) < sh_mul(kConstTol, d2)) ready |= bit;
            const double w0 = sh_sub(xx[i], xx[j]), w1 = sh_sub(xx[i + 1], xx[j + 1]), w2 = sh_sub(xx[i + 2], xx[j + 2]);
            const double scp = sh_add(sh_add(sh_mul(v0, w0), sh_mul(v1, w1)), sh_mul(v2, w2));
            const double wi = winv[a.i], wj = winv[a.j];
            const double corr = sh_div(diff, sh_mul(sh_mul(2.0, scp), sh_add(wi, wj)));
            // x(i) = x(i) + xxij*corr*winv(i);  x(j) = x(j) + (-xxij)*corr*winv(j)
            x[i] = sh_add(xi0, sh_mul(sh_mul(w0, corr), wi));
            x[i + 1] = sh_add(xi1, sh_mul(sh_mul(w1, corr), wi));
            x[i + 2] = sh_add(xi2, sh_mul(sh_mul(w2, corr), wi));
            x[j] = sh_add(xj0, sh_mul(sh_mul(-w0, corr), wj));
            x[j + 1] = sh_add(xj1, sh_mul(sh_mul(-w1, corr), wj));
            x[j + 2] = sh_add(xj2, sh_mul(sh_mul(-w2, corr), wj));
        }
        nits++;
        if (ready == all) break;
        if (nits >= kConstMaxIter) {   // die('shake failure'), bondene.f90:1143
            *failed = true;
            break;
        }
    }
    return nits;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(128)
k_shake(int nmol, const int *__restrict__ mol_first, const ShakePair *__restrict__ cij, const double *__restrict__ dist2,
        const double *__restrict__ winv, const double *__restrict__ xx, double *__restrict__ x,
        unsigned long long *__restrict__ iter_sum, int *__restrict__ failed) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmol) return;
    bool bad = false;
    const int nits = shake_molecule(mol_first[m], mol_first[m + 1], cij, dist2, winv, xx, x, &bad);
    if (bad) atomicExch(failed, 1);
    if (nits > 0) atomicAdd(iter_sum, (unsigned long long)nits);
}
#endif

}  // namespace qnb
