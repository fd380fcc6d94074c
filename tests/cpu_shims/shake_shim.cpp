// TEST INFRASTRUCTURE (CPU suite): runs the PRODUCT's shake_molecule() -- the very source the CUDA kernel k_shake
// executes per thread (q6_b200/csrc/qnb_shake.cuh, __host__ __device__) -- on the host, molecule by molecule.
// Built by tests/test_shake_cpu.py with g++ -ffp-contract=off.
#include "../../q6_b200/csrc/qnb_shake.cuh"

extern "C" int shake_shim(int nmol, const int *mol_first, const int *ij0 /*[2n] 0-based*/, const double *dist2,
                          const double *winv, const double *xx, double *x, long long *iter_sum) {
    int failed = 0;
    long long total = 0;
    for (int m = 0; m < nmol; m++) {
        bool bad = false;
        total += qnb::shake_molecule(mol_first[m], mol_first[m + 1], reinterpret_cast<const qnb::ShakePair *>(ij0), dist2, winv,
                                     xx, x, &bad);
        failed |= bad;
    }
    *iter_sum = total;
    return failed;
}
