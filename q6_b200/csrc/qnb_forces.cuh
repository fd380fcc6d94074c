// qnb_forces.cuh -- nonbonded force/energy kernels over the unit rows.
//
// Replaces nonbond_pp/_box, nonbond_pw/_box, nonbond_ww/_box, nonbond_ww_spc/_box (nbe, nbe_b, nbe_spc,
// nbe_spcb), nonbond_qp/_box, nonbond_qw/_box, nonbond_qw_spc/_box (nbe_qx, nbe_qspc), nonbond_qq,
// nonbond_qqp (nbe_qq) and lrf_taylor (nonbondene.f90:507-571, 4694-6077; nonbonded.f90:45-291).
//
// "d" is the reference's gradient array (globals.f90:347): d(i) -= vec*dv, d(j) += vec*dv with
// vec = x(j)-x(i) [+ periodic shift] and dv = (1/r) dV/dr.  Every kernel below accumulates the gradient of
// the atoms of ITS OWN unit only (full rows), so from atom a with partner b it always adds -vec(a->b)*dv.
// The kernels of one step run concurrently on several streams, hence the (uncontended) FP64 atomicAdd on grad.
#pragma once
#include "qnb_kernels.cuh"

namespace qnb {

#ifdef QNB_TRACE
// design experiment (tools/exp_trace.py): per-warp timestamps of the two persistent kernels
__device__ unsigned long long g_trace[2][8192][10];   // t0 c0 c1 t1 nchunk aux | tile loads, own / mirror / other-kind chunks
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define QTRACE(k, gw, i, v) do { if ((threadIdx.x & 31) == 0 && (gw) < 8192) g_trace[k][gw][i] = (v); } while (0)
#else
#define QTRACE(k, gw, i, v) do { } while (0)
#endif

// FP32 LJ parameters of a pair from per-type tables (precompute_set_values_*: simprep.f90:3326-3342)
template <bool GEOM>
__device__ __forceinline__ void lj_pair(const float *__restrict__ ljf, int cta, int ctb, int code, float &A, float &B) {
    const float2 pa = *reinterpret_cast<const float2 *>(ljf + (cta * 3 + code - 1) * 2);
    const float2 pb = *reinterpret_cast<const float2 *>(ljf + (ctb * 3 + code - 1) * 2);
    if (GEOM) {
        A = pa.x * pb.x;
        B = pa.y * pb.y;
    } else {
        float t = pa.x + pb.x;
        t = t * t;
        t = t * t * t;
        const float e = pa.y * pb.y;
        A = t * t * e;
        B = 2.0f * t * e;
    }
}

// One FP32 pair (nbe, nonbonded.f90:45-75): returns dv; accumulates LJ energy when asked.
template <bool LJ, bool ENERGY>
__device__ __forceinline__ float pair_f32(float dx, float dy, float dz, float qq, float A, float B, float &rinv_out,
                                          float &evdw) {
    const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    const float rinv = rsqrt_fast(r2);
    const float rinv2 = rinv * rinv;
    rinv_out = rinv;
    const float vel = qq * rinv;
    float t = -vel;
    if (LJ) {
        const float r6 = rinv2 * rinv2 * rinv2;
        const float va = A * r6 * r6, vb = B * r6;
        t = fmaf(6.0f, vb, fmaf(-12.0f, va, t));
        if (ENERGY) evdw += va - vb;
    }
    return rinv2 * t;
}

// FP64 Coulomb energy of one pair: elec / r with r from FP64 coordinates
__device__ __forceinline__ double coulomb_f64(double dx, double dy, double dz, double qq, float rinv_seed) {
    const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
    return qq * rsqrt_refine(r2, rinv_seed);
}
// FP64 Coulomb + LJ energies of one pair (clashing start structures carry LJ terms of 1e3..1e5 kcal/mol whose
// FP32 rounding alone would exceed 1e-6 of the net Vvdw)
__device__ __forceinline__ void energy_f64(double dx, double dy, double dz, double qq, double A, double B,
                                           float rinv_seed, double &eel, double &evdw) {
    const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
    const double y = rsqrt_refine(r2, rinv_seed);
    eel = fma(qq, y, eel);
    const double y2 = y * y, r6 = y2 * y2 * y2;
    evdw += fma(A * r6, r6, -B * r6);
}
// same, returning the two terms (callers add them under a select)
__device__ __forceinline__ void energy_terms_f64(double dx, double dy, double dz, double qq, double A, double B,
                                                 float rinv_seed, double &tel, double &tvdw) {
    const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
    const double y = rsqrt_refine(r2, rinv_seed);
    tel = qq * y;
    const double y2 = y * y, r6 = y2 * y2 * y2;
    tvdw = fma(A * r6, r6, -B * r6);
}
template <bool GEOM>
__device__ __forceinline__ void lj_pair_f64(const double *__restrict__ ljd, int cta, int ctb, int code, double &A, double &B) {
    const double2 pa = *reinterpret_cast<const double2 *>(ljd + (cta * 3 + code - 1) * 2);
    const double2 pb = *reinterpret_cast<const double2 *>(ljd + (ctb * 3 + code - 1) * 2);
    if (GEOM) {
        A = pa.x * pb.x;
        B = pa.y * pb.y;
    } else {
        double t = pa.x + pb.x;
        t = t * t;
        t = t * t * t;
        const double e = pa.y * pb.y;
        A = t * t * e;
        B = 2.0 * t * e;
    }
}

// ------------------------------------------------------------------------------------------------
// Water rows: ww (3x3 site tile) + the water side of pw.
//
// The rows are re-laid out at list-build time into CHUNKS of 32 entries (k_chunk_*: one descriptor {unit, kind}
// and 32 padded entries per chunk).  The kernel is persistent: every warp owns a contiguous range of chunks and
// runs a three-stage software pipeline over it -- descriptor+entry of chunk c+2 and partner coordinates of chunk
// c+1 are in flight while chunk c is computed -- because at 12k atoms the work is latency-bound (three dependent
// L2 round trips per chunk against ~500 issue cycles of arithmetic).  Gradients of the current water are kept in
// registers across its chunks and flushed (FP64 shuffle reduction + 9 atomicAdd) when the water changes.
constexpr int kChunkA = 0, kChunkB = 1;   // A: partner units of the own kind (own + mirror entries), B: the other kind
constexpr uint32_t kPadEntry = 0xffffffffu;

struct WaterTile {          // own molecule of the row being processed
    double o[3];            // row origin: own oxygen
    double sd[3][3];        // site offsets from the origin, FP64 (energies)
    float sf[3][3];         // same, FP32 (forces)
    float g[3][3];          // gradient accumulators
};

template <bool PBC, bool SPC>
__device__ __forceinline__ void ww_chunk(const Dev &D, WaterTile &T, bool own, bool any_own, bool valid,
                                         const double (&pj)[9], double &eel, double &evdw) {
    // Straight-line code for all 32 lanes: a padding lane computes on packed atom 0 and its dv is replaced by zero
    // (an early return of those lanes keeps the nine pairs from being scheduled together).
    // pj = x,y,z of the partner's three sites (SoA order: x0 x1 x2 y0 y1 y2 z0 z1 z2)
    double shx = 0, shy = 0, shz = 0;
    if (PBC) {
        // one shift per molecule pair from the O-O vector (nonbond_ww_spc_box L6016-6017)
        shx = pshift(T.o[0] - pj[0], D.box[0], D.inv_box[0]);
        shy = pshift(T.o[1] - pj[3], D.box[1], D.inv_box[1]);
        shz = pshift(T.o[2] - pj[6], D.box[2], D.inv_box[2]);
    }
    double ud[3][3];
    float uf[3][3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        ud[b][0] = (pj[b] - T.o[0]) + shx; ud[b][1] = (pj[3 + b] - T.o[1]) + shy; ud[b][2] = (pj[6 + b] - T.o[2]) + shz;
        uf[b][0] = (float)ud[b][0]; uf[b][1] = (float)ud[b][1]; uf[b][2] = (float)ud[b][2];
    }
    float seed[3][3];   // FP32 1/r of the nine site pairs: seeds of the FP64 energies
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const float dx = uf[b][0] - T.sf[a][0], dy = uf[b][1] - T.sf[a][1], dz = uf[b][2] - T.sf[a][2];
            float ev = 0.f, dv;
            const bool lj = !SPC || (a == 0 && b == 0);   // nonbond_ww_spc: only the first pair carries LJ
            if (lj) dv = pair_f32<true, false>(dx, dy, dz, D.wwQ[a * 3 + b], D.wwA[a * 3 + b], D.wwB[a * 3 + b], seed[a][b], ev);
            else dv = pair_f32<false, false>(dx, dy, dz, D.wwQ[a * 3 + b], 0.f, 0.f, seed[a][b], ev);
            dv = valid ? dv : 0.f;   // a select, not a product: the padding lane may hold r = 0
            T.g[a][0] = fmaf(-dx, dv, T.g[a][0]); T.g[a][1] = fmaf(-dy, dv, T.g[a][1]); T.g[a][2] = fmaf(-dz, dv, T.g[a][2]);
        }
    }
    // Energies once per pair, on the owner side, FP64.  ONE warp-uniform branch around all nine pairs (mirror-only chunks
    // skip it): nine independent dependency chains that the scheduler interleaves -- a branch per pair serialises them.
#ifdef QNB_EXP_OWNLANES
    if (own) {
#else
    if (any_own) {
#endif
        double ea[3] = {0.0, 0.0, 0.0}, ev = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const bool lj = !SPC || (a == 0 && b == 0);
                if (lj) energy_f64(ud[b][0] - T.sd[a][0], ud[b][1] - T.sd[a][1], ud[b][2] - T.sd[a][2], D.wwQd[a * 3 + b],
                                   D.wwAd[a * 3 + b], D.wwBd[a * 3 + b], seed[a][b], ea[a], ev);
                else ea[a] += coulomb_f64(ud[b][0] - T.sd[a][0], ud[b][1] - T.sd[a][1], ud[b][2] - T.sd[a][2], D.wwQd[a * 3 + b], seed[a][b]);
            }
        }
        if (own) { eel += (ea[0] + ea[1]) + ea[2]; evdw += ev; }
    }
}

// solute atom acting on the own water (pw, water side: gradient only)
template <bool PBC, bool GEOM>
__device__ __forceinline__ void wp_chunk(const Dev &D, WaterTile &T, bool valid, const double (&pj)[9], float qb, int ctb,
                                         uint32_t entry, const double *__restrict__ x, const int *__restrict__ pk_atom) {
    if (!valid) return;
    const int pb = (int)(entry & kIdMask);
    double ux = pj[0] - T.o[0], uy = pj[3] - T.o[1], uz = pj[6] - T.o[2];
    if (PBC) {
        if (D.any_atom) {
            // any-atom lists: the image found for the pair's reference atoms at list-build time (entry bits 24-29)
            ux += D.box[0] * img_comp(entry, 0); uy += D.box[1] * img_comp(entry, 1); uz += D.box[2] * img_comp(entry, 2);
        } else {
            // nonbond_pw_box: shift = boxlength*nint((x(solute switch)-x(water O))*inv_boxl), vec = x(j)-x(i)+shift
            // for i = solute atom, j = water atom; seen from the water: vec(a->b) = x(b)-x(a)-shift
            const int sw = D.g_switch[D.grp_of_atom[pk_atom[pb]]];
            ux -= pshift(x[3 * sw] - T.o[0], D.box[0], D.inv_box[0]);
            uy -= pshift(x[3 * sw + 1] - T.o[1], D.box[1], D.inv_box[1]);
            uz -= pshift(x[3 * sw + 2] - T.o[2], D.box[2], D.inv_box[2]);
        }
    }
    const float ufx = (float)ux, ufy = (float)uy, ufz = (float)uz;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int code = D.ljcode[ctb * D.nct + D.wct[a]];
        float A, B, rinv, ev = 0.f;
        lj_pair<GEOM>(D.ljf, D.wct[a], ctb, code, A, B);
        const float dx = ufx - T.sf[a][0], dy = ufy - T.sf[a][1], dz = ufz - T.sf[a][2];
        const float dv = pair_f32<true, false>(dx, dy, dz, D.wq[a] * qb, A, B, rinv, ev);
        T.g[a][0] = fmaf(-dx, dv, T.g[a][0]); T.g[a][1] = fmaf(-dy, dv, T.g[a][1]); T.g[a][2] = fmaf(-dz, dv, T.g[a][2]);
    }
}

#ifndef QNB_WATER_MINB
#define QNB_WATER_MINB 3   // measured on C2 and C5: 3 blocks (<=168 registers) beat 4 (spills) and 2
#endif
template <bool PBC, bool SPC, bool GEOM>
__global__ void __launch_bounds__(128, QNB_WATER_MINB)
k_water_force(Dev D, const double *__restrict__ x, const double *__restrict__ px, const double *__restrict__ py,
              const double *__restrict__ pz, const float *__restrict__ pk_q, const int *__restrict__ pk_ct,
              const int *__restrict__ pk_atom, const int *__restrict__ wstart, const int2 *__restrict__ cdesc,
              const uint32_t *__restrict__ crow, double *__restrict__ grad, double *__restrict__ Eslots, int nE,
              int want_energy) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int c0 = wstart[gw], c1 = wstart[gw + 1];   // this warp's share of the chunks (k_warp_starts)
    QTRACE(0, gw, 0, gtime()); QTRACE(0, gw, 1, (unsigned long long)clock64()); QTRACE(0, gw, 4, (unsigned long long)(c1 > c0 ? c1 - c0 : 0)); QTRACE(0, gw, 5, 0ull);
    if (c0 >= c1) return;
    WaterTile T;
    int cur_w = -1;
    double evdw = 0.0, eel = 0.0;
    const int slot = split_slot<9, 16>(lane);   // which gradient component this lane owns after the split reduction

    auto flush = [&]() {
        if (cur_w < 0) return;
        const int i0 = D.nat_solute + 3 * cur_w;
        float s[9];
#pragma unroll
        for (int k = 0; k < 9; k++) s[k] = T.g[k / 3][k % 3];
        const float mine = split_reduce<9, 16>(s, lane);
        if (slot >= 0) atomicAdd(&grad[3 * i0 + slot], (double)mine);
    };
    // stage A: descriptor + entry; stage B: partner coordinates.  Both are branch-free (padding entries fetch packed
    // atom 0, solute partners fetch three sites like waters do: the arrays are padded) so that the loads land directly
    // in the registers the NEXT iteration computes from: a conditional load costs a register move that waits for it.
    auto load_a = [&](int c, int2 &d, uint32_t &e) {
        d = cdesc[c];
        e = crow[(size_t)c * 32 + lane];
    };
    auto load_b = [&](uint32_t e, double (&pj)[9], float &qb, int &ctb) {
        const int p = (e == kPadEntry) ? 0 : (int)(e & kIdMask);
#pragma unroll
        for (int b = 0; b < 3; b++) { pj[b] = px[p + b]; pj[3 + b] = py[p + b]; pj[6 + b] = pz[p + b]; }
        qb = pk_q[p]; ctb = pk_ct[p];
    };
#ifdef QNB_TRACE
    unsigned long long tc_tile = 0, tc_own = 0, tc_mir = 0, tc_b = 0;
#endif
    auto compute = [&](const int2 &d, uint32_t e, const double (&pj)[9], float qb, int ctb) {
        if (d.x != cur_w) {
#ifdef QNB_TRACE
            tc_tile++;
#endif
            flush();
            cur_w = d.x;
            const int i0 = D.nat_solute + 3 * cur_w;
            T.o[0] = x[3 * i0]; T.o[1] = x[3 * i0 + 1]; T.o[2] = x[3 * i0 + 2];
#pragma unroll
            for (int a = 0; a < 3; a++) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    T.sd[a][k] = x[3 * (i0 + a) + k] - T.o[k];
                    T.sf[a][k] = (float)T.sd[a][k];
                    T.g[a][k] = 0.f;
                }
            }
        }
        const bool valid = e != kPadEntry;
#ifdef QNB_TRACE
        if (d.y == kChunkB) tc_b++;
#endif
        if (d.y == kChunkB) wp_chunk<PBC, GEOM>(D, T, valid, pj, qb, ctb, e, x, pk_atom);
        else {
            const bool own = valid && (e & kOwnerBit);
#ifdef QNB_TRACE
            if (__any_sync(kFull, own)) tc_own++; else tc_mir++;
#endif
            ww_chunk<PBC, SPC>(D, T, own, want_energy && __any_sync(kFull, own), valid, pj, eel, evdw);
        }
    };
    // chunk c computes from (d0,e0,P); chunk c+1 is (d1,e1) with its coordinates landing in R; (d2,e2) is chunk c+2.
    // The body is written out twice so that P and R swap roles without register moves.
    int2 d0, d1, d2;
    uint32_t e0, e1, e2;
    double P[9], R[9];
    float qP, qR;
    int ctP, ctR;
    const int clast = c1 - 1;
    load_a(c0, d0, e0);
    load_a(min(c0 + 1, clast), d1, e1);
    load_b(e0, P, qP, ctP);
    for (int c = c0; c < c1; c += 2) {
        load_a(min(c + 2, clast), d2, e2);
        load_b(e1, R, qR, ctR);
        compute(d0, e0, P, qP, ctP);
        if (c + 1 >= c1) break;
        load_a(min(c + 3, clast), d0, e0);
        load_b(e2, P, qP, ctP);
        compute(d1, e1, R, qR, ctR);
        d1 = d0; e1 = e0; d0 = d2; e0 = e2;   // chunk c+2 -> slot 0 (coordinates already in P), chunk c+3 -> slot 1
    }
    flush();
    QTRACE(0, gw, 2, (unsigned long long)clock64()); QTRACE(0, gw, 3, gtime());
#ifdef QNB_TRACE
    QTRACE(0, gw, 6, tc_tile); QTRACE(0, gw, 7, tc_own); QTRACE(0, gw, 8, tc_mir); QTRACE(0, gw, 9, tc_b);
#endif
    const double sv = warp_sum(evdw), se = warp_sum(eel);
    if (lane == 0) {
        double *E = Eslots + (size_t)(gw & (kESlots - 1)) * nE;
        atomicAdd(&E[QNB_E_WW_VDW], sv);
        atomicAdd(&E[QNB_E_WW_EL], se);
    }
}

// ------------------------------------------------------------------------------------------------
// Solute rows: pp + the solute (owner) side of pw.  Same persistent, software-pipelined structure as the water
// kernel; a chunk belongs to one (charge group, tile of up to four non-Q atoms) and the tile's gradient stays in
// registers across its chunks.  The per-type LJ tables live in shared memory.
constexpr int kITile = 4;

struct SoluteTile {
    int g, nt;
    int ai[kITile], cti[kITile];
    double o[3];
    double sd[kITile][3], qd[kITile];
    float sf[kITile][3], qf[kITile];
    float grad[kITile][3];
};

#ifndef QNB_SOLUTE_MINB
#define QNB_SOLUTE_MINB 2
#endif
template <bool PBC, bool GEOM>
__global__ void __launch_bounds__(128, QNB_SOLUTE_MINB)
k_solute_force(Dev D, const double *__restrict__ x, const double *__restrict__ px, const double *__restrict__ py,
               const double *__restrict__ pz, const float *__restrict__ pk_q, const double *__restrict__ pk_qd,
               const int *__restrict__ pk_ct, const int *__restrict__ pk_atom, const int *__restrict__ wstart,
               const int2 *__restrict__ cdesc, const uint32_t *__restrict__ crow, const uint16_t *__restrict__ cspec, double *__restrict__ grad,
               double *__restrict__ Eslots, int nE, int want_energy) {
    extern __shared__ unsigned char smem_raw[];
    // shared LJ tables: ljd [nct*6] doubles, ljf [nct*6] floats, ljcode [nct*nct] bytes
    double *s_ljd = reinterpret_cast<double *>(smem_raw);
    float *s_ljf = reinterpret_cast<float *>(s_ljd + D.nct * 6);
    unsigned char *s_code = reinterpret_cast<unsigned char *>(s_ljf + D.nct * 6);
    for (int k = threadIdx.x; k < D.nct * 6; k += blockDim.x) { s_ljd[k] = D.ljd[k]; s_ljf[k] = D.ljf[k]; }
    for (int k = threadIdx.x; k < D.nct * D.nct; k += blockDim.x) s_code[k] = D.ljcode[k];
#ifdef QNB_TRACE
    const unsigned long long tr0 = gtime();
#endif
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int c0 = wstart[gw], c1 = wstart[gw + 1];
#ifdef QNB_TRACE
    QTRACE(1, gw, 5, tr0);
#endif
    QTRACE(1, gw, 0, gtime()); QTRACE(1, gw, 1, (unsigned long long)clock64()); QTRACE(1, gw, 4, (unsigned long long)(c1 > c0 ? c1 - c0 : 0));
    if (c0 >= c1) return;
    SoluteTile T;
    T.g = -1; T.nt = 0;
    int cur_key = -1;
    double e_pp_el = 0.0, e_pw_el = 0.0, e_pp_vdw = 0.0, e_pw_vdw = 0.0;

    const int slot = split_slot<3 * kITile, 16>(lane);   // gradient component (3*tile atom + axis) this lane flushes
    auto flush = [&]() {
        if (cur_key < 0) return;
        float s[3 * kITile];
#pragma unroll
        for (int k = 0; k < 3 * kITile; k++) s[k] = T.grad[k / 3][k % 3];
        const float mine = split_reduce<3 * kITile, 16>(s, lane);
        const int t = slot / 3;
        int a = T.ai[0];
#pragma unroll
        for (int k = 1; k < kITile; k++) a = (t == k) ? T.ai[k] : a;
        if (slot >= 0 && t < T.nt) atomicAdd(&grad[3 * a + (slot - 3 * t)], (double)mine);
    };
    auto load_tile = [&](int g, int tile) {
        T.g = g;
        const int sw = D.g_switch[g];
        T.o[0] = x[3 * sw]; T.o[1] = x[3 * sw + 1]; T.o[2] = x[3 * sw + 2];   // row origin = switch atom
        const int k0 = D.nq_off[g] + tile * kITile;
        T.nt = min(kITile, D.nq_off[g + 1] - k0);
#pragma unroll
        for (int t = 0; t < kITile; t++) {
            T.ai[t] = -1; T.cti[t] = 0; T.qf[t] = 0.f; T.qd[t] = 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++) { T.sf[t][c] = 0.f; T.sd[t][c] = 0.0; T.grad[t][c] = 0.f; }
            if (t < T.nt) {
                const int a = D.nq_atoms[k0 + t];
                T.ai[t] = a; T.cti[t] = D.ctype[a]; T.qd[t] = D.crg[a]; T.qf[t] = (float)T.qd[t];
#pragma unroll
                for (int c = 0; c < 3; c++) { T.sd[t][c] = x[3 * a + c] - T.o[c]; T.sf[t][c] = (float)T.sd[t][c]; }
            }
        }
    };
    auto lj32 = [&](int cta, int ctb, int code, float &A, float &B) {
        const float ax = s_ljf[(cta * 3 + code - 1) * 2], ay = s_ljf[(cta * 3 + code - 1) * 2 + 1];
        const float bx = s_ljf[(ctb * 3 + code - 1) * 2], by = s_ljf[(ctb * 3 + code - 1) * 2 + 1];
        if (GEOM) { A = ax * bx; B = ay * by; }
        else { float t = ax + bx; t = t * t; t = t * t * t; const float e = ay * by; A = t * t * e; B = 2.0f * t * e; }
    };
    auto lj64 = [&](int cta, int ctb, int code, double &A, double &B) {
        const double ax = s_ljd[(cta * 3 + code - 1) * 2], ay = s_ljd[(cta * 3 + code - 1) * 2 + 1];
        const double bx = s_ljd[(ctb * 3 + code - 1) * 2], by = s_ljd[(ctb * 3 + code - 1) * 2 + 1];
        if (GEOM) { A = ax * bx; B = ay * by; }
        else { double t = ax + bx; t = t * t; t = t * t * t; const double e = ay * by; A = t * t * e; B = 2.0 * t * e; }
    };
    // pipeline stages (branch-free loads, see k_water_force)
    auto load_a = [&](int c, int2 &d, uint32_t &e) {
        d = cdesc[c];
        e = crow[(size_t)c * 32 + lane];
        d.y |= (int)cspec[(size_t)c * 32 + lane] << 16;   // per-lane special-pair codes ride in the descriptor's upper half
    };
    auto load_b = [&](uint32_t e, double (&pj)[9], float &qb, double &qbd, int &ctb) {
        const int p = (e == kPadEntry) ? 0 : (int)(e & kIdMask);
#pragma unroll
        for (int b = 0; b < 3; b++) { pj[b] = px[p + b]; pj[3 + b] = py[p + b]; pj[6 + b] = pz[p + b]; }
        qb = pk_q[p]; qbd = pk_qd[p]; ctb = pk_ct[p];
    };
#ifdef QNB_TRACE
    unsigned long long tc_tile = 0, tc_own = 0, tc_mir = 0, tc_b = 0;
#endif
    // One chunk.  Straight-line over the tile atoms: atoms beyond T.nt, padding lanes and excluded pairs are computed
    // and then dropped by a select, so that the tile's pairs (and afterwards their FP64 energies) form independent
    // dependency chains the scheduler can interleave.
    auto compute = [&](const int2 &d, uint32_t e, const double (&pj)[9], float qb, double qbd, int ctb) {
        const int tile = (d.y >> 8) & 0xff;
        const int key = d.x * 256 + tile;   // (group, tile)
#ifdef QNB_TRACE
        if (key != cur_key) tc_tile++;
        if ((d.y & 0xff) != kChunkA) tc_b++;
        else if (__any_sync(kFull, e != kPadEntry && (e & kOwnerBit) != 0)) tc_own++; else tc_mir++;
#endif
        if (key != cur_key) { flush(); load_tile(d.x, tile); cur_key = key; }
        const bool valid = e != kPadEntry;
        if ((d.y & 0xff) == kChunkA) {
            // ---- solute-solute partner atom
            const int pb = (int)(e & kIdMask);
            const bool own = valid && (e & kOwnerBit) != 0;
            double ux = pj[0] - T.o[0], uy = pj[3] - T.o[1], uz = pj[6] - T.o[2];
            if (PBC) {
                if (D.any_atom) {
                    ux += D.box[0] * img_comp(e, 0); uy += D.box[1] * img_comp(e, 1); uz += D.box[2] * img_comp(e, 2);
                } else if (valid) {
                    // nonbond_pp_box L4791-4801: shift = boxlength*nint((x(sw_i)-x(sw_j))*inv_boxl)
                    const int swb = D.g_switch[D.grp_of_atom[pk_atom[pb]]];
                    ux += pshift(T.o[0] - x[3 * swb], D.box[0], D.inv_box[0]);
                    uy += pshift(T.o[1] - x[3 * swb + 1], D.box[1], D.inv_box[1]);
                    uz += pshift(T.o[2] - x[3 * swb + 2], D.box[2], D.inv_box[2]);
                }
            }
            const float ufx = (float)ux, ufy = (float)uy, ufz = (float)uz;
            // special pairs (exclusions, 1-4 neighbours, own group) were resolved per tile atom when the chunks were
            // laid out (k_chunk_fill): 3 bits per tile atom = skip | 1-4 | energy on the other side
            const unsigned spec = (unsigned)d.y >> 16;
            int code[kITile];
            bool keep[kITile], lower[kITile], i14[kITile];   // evaluated at all; energy side inside one group (i<j, L1874)
#pragma unroll
            for (int t = 0; t < kITile; t++) {
                const unsigned sb = (spec >> (3 * t)) & 7u;
                i14[t] = (sb & 2u) != 0;
                code[t] = i14[t] ? 3 : s_code[T.cti[t] * D.nct + ctb];
                keep[t] = valid && t < T.nt && !(sb & 1u);
                lower[t] = !(sb & 4u);
            }
            float seed[kITile];
#pragma unroll
            for (int t = 0; t < kITile; t++) {
                float A, B, ev = 0.f;
                lj32(T.cti[t], ctb, code[t], A, B);
                const float qq = i14[t] ? T.qf[t] * qb * D.el14f : T.qf[t] * qb;
                const float dx = ufx - T.sf[t][0], dy = ufy - T.sf[t][1], dz = ufz - T.sf[t][2];
                float dv = pair_f32<true, false>(dx, dy, dz, qq, A, B, seed[t], ev);
                dv = keep[t] ? dv : 0.f;
                T.grad[t][0] = fmaf(-dx, dv, T.grad[t][0]); T.grad[t][1] = fmaf(-dy, dv, T.grad[t][1]);
                T.grad[t][2] = fmaf(-dz, dv, T.grad[t][2]);
            }
            // energy once per pair: on the owner side
            if (want_energy && __any_sync(kFull, own)) {
#pragma unroll
                for (int t = 0; t < kITile; t++) {
                    const double qqd = i14[t] ? T.qd[t] * qbd * D.el14 : T.qd[t] * qbd;
                    double Ad, Bd, tel, tvdw;
                    lj64(T.cti[t], ctb, code[t], Ad, Bd);
                    energy_terms_f64(ux - T.sd[t][0], uy - T.sd[t][1], uz - T.sd[t][2], qqd, Ad, Bd, seed[t], tel, tvdw);
                    const bool add = own && keep[t] && lower[t];
                    e_pp_el += add ? tel : 0.0; e_pp_vdw += add ? tvdw : 0.0;
                }
            }
        } else {
            // ---- solute-water: this side owns the pair, all three water atoms with full LJ (nbe)
            double shx = 0, shy = 0, shz = 0;
            if (PBC) {
                if (D.any_atom) {
                    shx = D.box[0] * img_comp(e, 0); shy = D.box[1] * img_comp(e, 1); shz = D.box[2] * img_comp(e, 2);
                } else {
                    shx = pshift(T.o[0] - pj[0], D.box[0], D.inv_box[0]);
                    shy = pshift(T.o[1] - pj[3], D.box[1], D.inv_box[1]);
                    shz = pshift(T.o[2] - pj[6], D.box[2], D.inv_box[2]);
                }
            }
#pragma unroll
            for (int sw = 0; sw < 3; sw++) {
                const double ux = (pj[sw] - T.o[0]) + shx, uy = (pj[3 + sw] - T.o[1]) + shy, uz = (pj[6 + sw] - T.o[2]) + shz;
                const float ufx = (float)ux, ufy = (float)uy, ufz = (float)uz;
                const int ctw = D.wct[sw];
                float seed[kITile];
                int code[kITile];
#pragma unroll
                for (int t = 0; t < kITile; t++) {
                    code[t] = s_code[T.cti[t] * D.nct + ctw];
                    float A, B, ev = 0.f;
                    lj32(T.cti[t], ctw, code[t], A, B);
                    const float dx = ufx - T.sf[t][0], dy = ufy - T.sf[t][1], dz = ufz - T.sf[t][2];
                    float dv = pair_f32<true, false>(dx, dy, dz, T.qf[t] * D.wq[sw], A, B, seed[t], ev);
                    dv = (valid && t < T.nt) ? dv : 0.f;
                    T.grad[t][0] = fmaf(-dx, dv, T.grad[t][0]); T.grad[t][1] = fmaf(-dy, dv, T.grad[t][1]);
                    T.grad[t][2] = fmaf(-dz, dv, T.grad[t][2]);
                }
                if (want_energy)
#pragma unroll
                for (int t = 0; t < kITile; t++) {
                    double Ad, Bd, tel, tvdw;
                    lj64(T.cti[t], ctw, code[t], Ad, Bd);
                    energy_terms_f64(ux - T.sd[t][0], uy - T.sd[t][1], uz - T.sd[t][2], T.qd[t] * D.wqd[sw], Ad, Bd, seed[t], tel, tvdw);
                    const bool add = valid && t < T.nt;
                    e_pw_el += add ? tel : 0.0; e_pw_vdw += add ? tvdw : 0.0;
                }
            }
        }
    };
    int2 d0, d1, d2;
    uint32_t e0, e1, e2;
    double P[9], R[9], qdP, qdR;
    float qP, qR;
    int ctP, ctR;
    const int clast = c1 - 1;
    load_a(c0, d0, e0);
    load_a(min(c0 + 1, clast), d1, e1);
    load_b(e0, P, qP, qdP, ctP);
    for (int c = c0; c < c1; c += 2) {
        load_a(min(c + 2, clast), d2, e2);
        load_b(e1, R, qR, qdR, ctR);
        compute(d0, e0, P, qP, qdP, ctP);
        if (c + 1 >= c1) break;
        load_a(min(c + 3, clast), d0, e0);
        load_b(e2, P, qP, qdP, ctP);
        compute(d1, e1, R, qR, qdR, ctR);
        d1 = d0; e1 = e0; d0 = d2; e0 = e2;
    }
    flush();
    QTRACE(1, gw, 2, (unsigned long long)clock64()); QTRACE(1, gw, 3, gtime());
#ifdef QNB_TRACE
    QTRACE(1, gw, 6, tc_tile); QTRACE(1, gw, 7, tc_own); QTRACE(1, gw, 8, tc_mir); QTRACE(1, gw, 9, tc_b);
#endif
    const double s1 = warp_sum(e_pp_el), s2 = warp_sum(e_pp_vdw), s3 = warp_sum(e_pw_el), s4 = warp_sum(e_pw_vdw);
    if (lane == 0) {
        double *E = Eslots + (size_t)(gw & (kESlots - 1)) * nE;
        atomicAdd(&E[QNB_E_PP_EL], s1); atomicAdd(&E[QNB_E_PP_VDW], s2);
        atomicAdd(&E[QNB_E_PW_EL], s3); atomicAdd(&E[QNB_E_PW_VDW], s4);
    }
}

// ------------------------------------------------------------------------------------------------
// Q-atom kernels, FP64 throughout (per-state energies are FEP observables), written without divisions:
// 1/r from an FP32 rsqrt seed + Halley step, 1/(r^6+alpha) from an FP32 reciprocal seed + two Newton steps.
struct QxOut { double vel, vvdw, dv; };
__device__ __forceinline__ double rcp_refine(double d) {
    double z = (double)__frcp_rn((float)d);
    z = z * fma(-d, z, 2.0);
    z = z * fma(-d, z, 2.0);
    return z;
}
// nbe_qx (nonbonded.f90:200-222) for one state; r2 = |vec|^2, rinv = 1/r
__device__ __forceinline__ QxOut qx_eval(double r2, double rinv, const QPar4 &p, double lambda) {
    QxOut o;
    const double r2inv = rinv * rinv;
    const double r6_hc = r2 * r2 * r2;                 // 1/dist%r6 = r^6
    const double r6s = rcp_refine(r6_hc + p.score);    // softcore: 1/(r^6 + alpha)
    const double r12 = r6s * r6s;
    o.vel = p.el * rinv;
    const double va = p.A * r12, vb = p.B * r6s;
    o.vvdw = va - vb;
    o.dv = r2inv * (-o.vel - (12.0 * va - 6.0 * vb) * r6s * r6_hc) * lambda;
    return o;
}

// Partner site p of the Q lists: [0,nqp) listed solute atoms, then 3 sites per listed water.
struct QSite { int atom; int site; bool water; };
__device__ __forceinline__ QSite q_site(const Dev &D, int p, int nqp, const int *qp_list, const int *qw_list) {
    QSite s;
    if (p < nqp) { s.atom = qp_list[p]; s.site = 0; s.water = false; }
    else { const int k = p - nqp; s.site = k % 3; s.atom = D.nat_solute + 3 * qw_list[k / 3] + s.site; s.water = true; }
    return s;
}
// periodic shift of the Q -> partner vector (vec = shift - (x(i) - x(j)))
template <bool PBC>
__device__ __forceinline__ void q_shift(const Dev &D, const double *__restrict__ x, const QSite &s,
                                        const int *__restrict__ qp_shift_atom, double sh[3]) {
    sh[0] = sh[1] = sh[2] = 0.0;
    if (!PBC) return;
    const int qs = D.qswitch0;
    if (!s.water) {
        // nonbond_qp_box L5275-5278: boxlength*nint((x(ia)-x(qswitch))*inv_boxl), ia = the group's first atom inside Rcq;
        // no registered pair (Rq<0) -> zero shift
        const int ref = qp_shift_atom[D.grp_of_atom[s.atom]];
        if (ref >= 0)
            for (int c = 0; c < 3; c++) sh[c] = pshift(x[3 * ref + c] - x[3 * qs + c], D.box[c], D.inv_box[c]);
    } else {
        // spc: one shift per molecule from x(qswitch)-x(O) (L5818-5820); general: per atom j (L5470-5472)
        const int r = D.spc_water ? s.atom - s.site : s.atom;
        for (int c = 0; c < 3; c++) sh[c] = pshift(x[3 * qs + c] - x[3 * r + c], D.box[c], D.inv_box[c]);
    }
}

// Gradient on the PARTNER atoms (nonbond_qp/_box, nonbond_qw(_spc)/_box seen from atom j): one thread per
// partner site, all Q-atoms looped from shared memory.  No energies leave this kernel, so it follows the rule of
// the other gradient paths: FP64 only to form coordinates relative to a local origin (the Q switch atom; periodic
// shift folded in), FP32 pair arithmetic with FP32 parameter tables.
template <bool PBC>
struct QPartnerBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(const Dev &D, const double *__restrict__ x, const double *__restrict__ lambda, int nqp, const int *__restrict__ qp_list, const int *__restrict__ qp_shift_atom, int nqw, const int *__restrict__ qw_list, double *__restrict__ grad, const int BX, const int NBX, const int BY, const int NBY) {
    extern __shared__ float shq[];
    float *xq = shq;                 // [nqat][3] relative to the origin
    float *lam = shq + 3 * D.nqat;   // [nstates]
    const int org = D.iqseq[0];
    const double ox = x[3 * org], oy = x[3 * org + 1], oz = x[3 * org + 2];
    for (int k = threadIdx.x; k < D.nqat; k += blockDim.x) {
        const int a = D.iqseq[k];
        xq[3 * k] = (float)(x[3 * a] - ox); xq[3 * k + 1] = (float)(x[3 * a + 1] - oy); xq[3 * k + 2] = (float)(x[3 * a + 2] - oz);
    }
    for (int k = threadIdx.x; k < D.nstates; k += blockDim.x) lam[k] = (float)lambda[k];
    __syncthreads();
    const int p = BX * blockDim.x + threadIdx.x;
    if (p >= nqp + 3 * nqw) return;
    const int nst = D.nstates;
    const QSite s = q_site(D, p, nqp, qp_list, qw_list);
    double shf[3];
    q_shift<PBC>(D, x, s, qp_shift_atom, shf);
    // vec = shift - (x(i) - x(j)) = (x(j) + shift - origin) - (x(i) - origin)
    const float jx = (float)((x[3 * s.atom] - ox) + shf[0]), jy = (float)((x[3 * s.atom + 1] - oy) + shf[1]),
                jz = (float)((x[3 * s.atom + 2] - oz) + shf[2]);
    const bool coul_only = s.water && D.spc_water && s.site > 0;   // nbe_qspc
    float gx = 0.f, gy = 0.f, gz = 0.f;
    // the Q-atoms are dealt to NBY blocks so that the few thousand partner sites still fill the machine
    const int qchunk = (D.nqat + NBY - 1) / NBY;
    const int q0 = BY * qchunk, q1 = min(D.nqat, q0 + qchunk);
#pragma unroll 2
    for (int q = q0; q < q1; q++) {
        const float vx = jx - xq[3 * q], vy = jy - xq[3 * q + 1], vz = jz - xq[3 * q + 2];
        const float r2 = fmaf(vx, vx, fmaf(vy, vy, vz * vz)), rinv = rsqrt_fast(r2), r2inv = rinv * rinv;
        const float r6_hc = r2 * r2 * r2;
        float dv = 0.f;
        for (int st = 0; st < nst; st++) {
            const float4 pr = s.water ? D.qw_tabf[(size_t)(q * nst + st) * 3 + s.site]
                                      : D.qp_tabf[(size_t)(q * nst + st) * D.nat_solute + s.atom];   // A B el score
            const float vel = pr.z * rinv;
            float t = -vel;
            if (!coul_only) {
                // nbe_qx (nonbonded.f90:200-222): softcore 1/(r^6 + alpha)
                const float r6s = __fdividef(1.0f, r6_hc + pr.w);
                const float va = pr.x * r6s * r6s, vb = pr.y * r6s;
                t -= (12.0f * va - 6.0f * vb) * r6s * r6_hc;
            }
            dv = fmaf(r2inv * t, lam[st], dv);
        }
        gx = fmaf(vx, dv, gx); gy = fmaf(vy, dv, gy); gz = fmaf(vz, dv, gz);   // d(j) += vec*dv
    }
    atomicAdd(&grad[3 * s.atom], (double)gx); atomicAdd(&grad[3 * s.atom + 1], (double)gy); atomicAdd(&grad[3 * s.atom + 2], (double)gz);
}
};
template <bool PBC>
__global__ void __launch_bounds__(128)
k_q_partner(Dev D, const double *__restrict__ x, const double *__restrict__ lambda, int nqp,
            const int *__restrict__ qp_list, const int *__restrict__ qp_shift_atom, int nqw,
            const int *__restrict__ qw_list, double *__restrict__ grad) {
    QPartnerBody<PBC>::run(D, x, lambda, nqp, qp_list, qp_shift_atom, nqw, qw_list, grad, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// Gradient on the Q-atoms and the per-state energies EQ(:)%qp, EQ(:)%qw.  Block (q, slice): the partner sites
// are dealt round-robin to gridDim.y slices so that a few dozen Q-atoms still fill the machine.
template <bool PBC, int NS>
struct QAtomBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(const Dev &D, const double *__restrict__ x, const double *__restrict__ lambda, int nqp, const int *__restrict__ qp_list, const int *__restrict__ qp_shift_atom, int nqw, const int *__restrict__ qw_list, double *__restrict__ grad, double *__restrict__ Eslots, int nE, const int BX, const int NBX, const int BY, const int NBY) {
    __shared__ double red[4][3 + 4 * NS];
    double *EQ = Eslots + (size_t)((BX + BY) & (kESlots - 1)) * nE + QNB_E_COUNT;
    const int q = BX, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nst = D.nstates;
    const int iat = D.iqseq[q];
    const double qx = x[3 * iat], qy = x[3 * iat + 1], qz = x[3 * iat + 2];
    double lam[NS], eel[NS], evdw[NS], wel[NS], wvdw[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) { lam[s] = s < nst ? lambda[s] : 0.0; eel[s] = evdw[s] = wel[s] = wvdw[s] = 0.0; }
    double gx = 0, gy = 0, gz = 0;
    const int nsite = nqp + 3 * nqw;
    for (int p = BY * blockDim.x + tid; p < nsite; p += NBY * blockDim.x) {
        const QSite s = q_site(D, p, nqp, qp_list, qw_list);
        double shf[3];
        q_shift<PBC>(D, x, s, qp_shift_atom, shf);
        const double vx = shf[0] - (qx - x[3 * s.atom]), vy = shf[1] - (qy - x[3 * s.atom + 1]), vz = shf[2] - (qz - x[3 * s.atom + 2]);
        const double r2 = vx * vx + vy * vy + vz * vz, rinv = rinv_f64(r2);
        const bool coul_only = s.water && D.spc_water && s.site > 0;
        double dv = 0;
#pragma unroll
        for (int st = 0; st < NS; st++)
            if (st < nst) {
                const QPar4 pr = s.water ? D.qw_tab[(size_t)(q * nst + st) * 3 + s.site]
                                         : D.qp_tab[(size_t)(q * nst + st) * D.nat_solute + s.atom];
                if (coul_only) {
                    const double vel = pr.el * rinv;
                    wel[st] += vel; dv += -(rinv * rinv) * vel * lam[st];
                } else {
                    const QxOut o = qx_eval(r2, rinv, pr, lam[st]);
                    if (s.water) { wel[st] += o.vel; wvdw[st] += o.vvdw; } else { eel[st] += o.vel; evdw[st] += o.vvdw; }
                    dv += o.dv;
                }
            }
        gx -= vx * dv; gy -= vy * dv; gz -= vz * dv;   // d(i) -= vec*dv
    }
    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
#pragma unroll
    for (int s = 0; s < NS; s++)
        if (s < nst) { eel[s] = warp_sum(eel[s]); evdw[s] = warp_sum(evdw[s]); wel[s] = warp_sum(wel[s]); wvdw[s] = warp_sum(wvdw[s]); }
    if (lane == 0) {
        red[wid][0] = gx; red[wid][1] = gy; red[wid][2] = gz;
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (s < nst) { red[wid][3 + 4 * s] = eel[s]; red[wid][4 + 4 * s] = evdw[s]; red[wid][5 + 4 * s] = wel[s]; red[wid][6 + 4 * s] = wvdw[s]; }
    }
    __syncthreads();
    if (tid < 3 + 4 * nst) {
        const int nw = blockDim.x >> 5;
        double a = 0;
        for (int k = 0; k < nw; k++) a += red[k][tid];
        if (tid < 3) atomicAdd(&grad[3 * iat + tid], a);
        else {
            const int s = (tid - 3) >> 2, c = (tid - 3) & 3;
            atomicAdd(&EQ[QNB_EQ_STRIDE * s + 2 + c], a);
        }
    }
}
};
template <bool PBC, int NS>
__global__ void __launch_bounds__(128)
k_q_atom(Dev D, const double *__restrict__ x, const double *__restrict__ lambda, int nqp,
         const int *__restrict__ qp_list, const int *__restrict__ qp_shift_atom, int nqw,
         const int *__restrict__ qw_list, double *__restrict__ grad, double *__restrict__ Eslots, int nE) {
    QAtomBody<PBC, NS>::run(D, x, lambda, nqp, qp_list, qp_shift_atom, nqw, qw_list, grad, Eslots, nE, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// Static lists nbqq / nbqqp (nonbond_qq L5013, nonbond_qqp L5085): one thread per (pair,state) entry.
struct QStatic { int i, j, state, soft; QPar4 p; };
struct QqStaticBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(int n, int nqq, const QStatic *__restrict__ lst, const double *__restrict__ x, const double *__restrict__ lambda, double *__restrict__ grad, double *__restrict__ Eslots, int nE, const int BX, const int NBX, const int BY, const int NBY) {
    const int k = BX * blockDim.x + threadIdx.x;
    double *EQ = Eslots + (size_t)(k & (kESlots - 1)) * nE + QNB_E_COUNT;
    if (k >= n) return;
    const QStatic e = lst[k];
    const double vx = x[3 * e.j] - x[3 * e.i], vy = x[3 * e.j + 1] - x[3 * e.i + 1], vz = x[3 * e.j + 2] - x[3 * e.i + 2];
    const double r2 = vx * vx + vy * vy + vz * vz, rinv = rinv_f64(r2), r2inv = rinv * rinv;
    const double lam = lambda[e.state];
    double vel, vvdw, dv;
    if (e.soft) {
        // nbe_qq soft pair: V_a = A*exp(-B*r), code "-nb%vdWB/r" with r holding 1/r (nonbonded.f90:140-143)
        const double r = r2 * rinv;
        vel = e.p.el * rinv;
        const double va = e.p.A * exp(-e.p.B * r);
        vvdw = va;
        dv = r2inv * (-vel - e.p.B * va * r) * lam;
    } else {
        const QxOut o = qx_eval(r2, rinv, e.p, lam);
        vel = o.vel; vvdw = o.vvdw; dv = o.dv;
    }
    atomicAdd(&grad[3 * e.i], -vx * dv); atomicAdd(&grad[3 * e.i + 1], -vy * dv); atomicAdd(&grad[3 * e.i + 2], -vz * dv);
    atomicAdd(&grad[3 * e.j], vx * dv); atomicAdd(&grad[3 * e.j + 1], vy * dv); atomicAdd(&grad[3 * e.j + 2], vz * dv);
    const int o = (k < nqq) ? 0 : 2;   // nbqq -> EQ%qq, nbqqp -> EQ%qp (potene.f90:176-177)
    atomicAdd(&EQ[QNB_EQ_STRIDE * e.state + o], vel);
    atomicAdd(&EQ[QNB_EQ_STRIDE * e.state + o + 1], vvdw);
}
};
__global__ void
k_qq_static(int n, int nqq, const QStatic *__restrict__ lst, const double *__restrict__ x,
                            const double *__restrict__ lambda, double *__restrict__ grad, double *__restrict__ Eslots, int nE) {
    QqStaticBody::run(n, nqq, lst, x, lambda, grad, Eslots, nE, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// lrf_taylor (nonbondene.f90:507-571)
struct LrfTaylorBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(const Dev &D, const double *__restrict__ x, const double *__restrict__ lrf, double *__restrict__ grad, double *__restrict__ Eslots, int nE, const int BX, const int NBX, const int BY, const int NBY) {
    const int i = BX * blockDim.x + threadIdx.x;
    double *E = Eslots + (size_t)(BX & (kESlots - 1)) * nE;
    double e = 0.0;
    if (i < D.natom && i + 1 >= D.at_s && i + 1 <= D.at_e && !D.is_q[i] && (D.use_PBC || !D.excl[i])) {
        const double *l = lrf + (size_t)QNB_LRF_STRIDE * D.grp_of_atom[i];
        const double dx = l[0] - x[3 * i], dy = l[1] - x[3 * i + 1], dz = l[2] - x[3 * i + 2];
        const double *p1 = l + 4, *p2 = l + 7, *p3 = l + 16;
        // potential
        const double t0 = dx * p2[0] + dy * p2[1] + dz * p2[2];
        const double t1 = dx * p2[3] + dy * p2[4] + dz * p2[5];
        const double t2 = dx * p2[6] + dy * p2[7] + dz * p2[8];
        const double V = l[3] + (dx * p1[0] + dy * p1[1] + dz * p1[2]) + 0.5 * (dx * t0 + dy * t1 + dz * t2);
        const double q = D.crg[i];
        e = 0.5 * q * V;
        // field
        double df[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const double *r = p3 + 9 * a;
            const double u0 = dx * r[0] + dy * r[1] + dz * r[2];
            const double u1 = dx * r[3] + dy * r[4] + dz * r[5];
            const double u2 = dx * r[6] + dy * r[7] + dz * r[8];
            const double d2 = (a == 0) ? t0 : (a == 1) ? t1 : t2;
            df[a] = p1[a] + d2 + 0.5 * (dx * u0 + dy * u1 + dz * u2);
        }
        atomicAdd(&grad[3 * i], -df[0] * q); atomicAdd(&grad[3 * i + 1], -df[1] * q); atomicAdd(&grad[3 * i + 2], -df[2] * q);
    }
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(&E[QNB_E_LRF], e);
}
};
__global__ void
k_lrf_taylor(Dev D, const double *__restrict__ x, const double *__restrict__ lrf,
                             double *__restrict__ grad, double *__restrict__ Eslots, int nE) {
    LrfTaylorBody::run(D, x, lrf, grad, Eslots, nE, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// ------------------------------------------------------------------------------------------------
// Spherical-boundary solvent restraints, FP64 throughout.
// rst = [E solvent_radial, E water_pol, theta sums per shell (QNB_MAX_SHELLS), n_insh per shell (QNB_MAX_SHELLS),
//        QNB_MAX_SHELLS int32 append counters of the shell lists]: part of the step's output buffer, cleared with it
struct RstPar {
    double xw[3], rsurf /* rwat - shift */, fk, Dwmz, awmz, fkwpol;
    int wpol, nsh;
    double rout[QNB_MAX_SHELLS], cstb[QNB_MAX_SHELLS], tcorr[QNB_MAX_SHELLS];
    double rin_last;   // rout(nsh) - dr(nsh): inner edge of the innermost shell
};
constexpr int kRstOut = 2 + 2 * QNB_MAX_SHELLS + QNB_MAX_SHELLS / 2;   // + the shell counters (int32), cleared with the buffer

// water-molecule geometry of watpol (L6583-6620): unit dipole direction rmu, its length rm, unit radial vector rcu, rc
struct WpGeom { double rmu[3], rcu[3], rm, rc, scp; };
__device__ __forceinline__ WpGeom wp_geom(const double *__restrict__ x, int i, const RstPar &P) {
    WpGeom g;
#pragma unroll
    for (int c = 0; c < 3; c++) g.rmu[c] = (x[3 * (i + 1) + c] + x[3 * (i + 2) + c]) - x[3 * i + c] * 2.0;
    g.rm = sqrt(g.rmu[0] * g.rmu[0] + g.rmu[1] * g.rmu[1] + g.rmu[2] * g.rmu[2]);
#pragma unroll
    for (int c = 0; c < 3; c++) { g.rmu[c] /= g.rm; g.rcu[c] = x[3 * i + c] - P.xw[c]; }
    g.rc = sqrt(g.rcu[0] * g.rcu[0] + g.rcu[1] * g.rcu[1] + g.rcu[2] * g.rcu[2]);
#pragma unroll
    for (int c = 0; c < 3; c++) g.rcu[c] /= g.rc;
    g.scp = fmin(1.0, fmax(-1.0, g.rmu[0] * g.rcu[0] + g.rmu[1] * g.rcu[1] + g.rmu[2] * g.rcu[2]));
    return g;
}

// restrain_solvent (nonbondene.f90:6486-6509) and the first loop of watpol (L6565-6625): theta and shell membership
__global__ void k_rst_theta(Dev D, RstPar P, const double *__restrict__ x, double *__restrict__ grad,
                            double *__restrict__ rst, double *__restrict__ theta, int *__restrict__ shell_n,
                            int *__restrict__ shell_list, double *__restrict__ shell_theta) {
    const int iw = blockIdx.x * blockDim.x + threadIdx.x;
    double erst = 0.0;
    if (iw < D.nwat) {
        const int i = D.nat_solute + 3 * iw;
        if (!D.excl[i]) {
            const double vx = x[3 * i] - P.xw[0], vy = x[3 * i + 1] - P.xw[1], vz = x[3 * i + 2] - P.xw[2];
            const double b = sqrt(vx * vx + vy * vy + vz * vz);
            const double db = b - P.rsurf;
            double dv = 0.0;
            if (db > 0.0) {
                erst = 0.5 * P.fk * db * db - P.Dwmz;
                dv = P.fk * db / b;
            } else if (b > 0.0) {
                const double fexp = exp(P.awmz * db);
                erst = P.Dwmz * (fexp * fexp - 2.0 * fexp);
                dv = -2.0 * P.Dwmz * P.awmz * (fexp - fexp * fexp) / b;
            }
            atomicAdd(&grad[3 * i], vx * dv); atomicAdd(&grad[3 * i + 1], vy * dv); atomicAdd(&grad[3 * i + 2], vz * dv);
            if (P.wpol) {
                const WpGeom g = wp_geom(x, i, P);
                theta[iw] = acos(g.scp);
                if (g.rc > P.rin_last) {
                    int is = P.nsh;                       // do is = nwpolr_shell, 2, -1; if (rc <= rout(is)) exit
                    while (is >= 2 && !(g.rc <= P.rout[is - 1])) is--;
                    const int pos = atomicAdd(&shell_n[is - 1], 1);
                    shell_list[(size_t)(is - 1) * D.nwat + pos] = iw;
                    shell_theta[(size_t)(is - 1) * D.nwat + pos] = theta[iw];   // compact copy for the ranking
                }
            }
        } else if (P.wpol) theta[iw] = 0.0;
    }
    erst = warp_sum(erst);
    if ((threadIdx.x & 31) == 0 && erst != 0.0) atomicAdd(&rst[0], erst);
}

// second part of watpol (L6627-6742): rank of the molecule inside its shell by (theta, molecule number) -- the order the
// reference's selection sort produces --, target angle of that rank, energy and gradient.  blockIdx.y = shell.
__global__ void k_rst_watpol(Dev D, RstPar P, const double *__restrict__ x, double *__restrict__ grad,
                             double *__restrict__ rst, const double *__restrict__ theta,
                             const int *__restrict__ shell_n, const int *__restrict__ shell_list,
                             const double *__restrict__ shell_theta) {
    // one WARP per shell member (grid-stride over the members): its lanes share the n comparisons of the ranking (a
    // thread per member walks them serially: 600 dependent iterations on a handful of warps took 43 us)
    const int is = blockIdx.y, n = shell_n[is];
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, nwarp = gridDim.x * wpb;
    const int *lst = shell_list + (size_t)is * D.nwat;
    const double *sth = shell_theta + (size_t)is * D.nwat;
    double e = 0.0, tsum = 0.0;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += nwarp) {
        const int iw = lst[j];
        const double th = sth[j];
        int cnt = 0;
        for (int k = lane; k < n; k += 32) {
            const double t = sth[k];
            cnt += (int)(t < th) | ((int)(t == th) & (int)(lst[k] < iw));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
        if (lane != 0) continue;
        const int il = cnt + 1;
        const double pi = 3.14159265358979323846;
        const double arg = 1.0 + (1.0 - 2.0 * (double)il) / (double)n;
        double theta0 = acos(arg);
        theta0 = theta0 - 3.0 * sin(theta0) * P.cstb[is] / 2.0;
        theta0 = fmin(pi, fmax(0.0, theta0));
        const double dth = th - theta0 + P.tcorr[is];
        e += 0.5 * P.fkwpol * dth * dth;
        tsum += th;
        const double dv = P.fkwpol * dth;
        const int i = D.nat_solute + 3 * iw;
        const WpGeom g = wp_geom(x, i, P);
        double f0 = sin(acos(g.scp));
        if (fabs(f0) < 1.0e-10) f0 = 1.0e-10;   // QREAL_EPS (sizes.f90:50)
        f0 = dv * (-1.0 / f0);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double f1 = ((g.rcu[c] - g.rmu[c] * g.scp) * (-2.0)) / g.rm;
            const double f3 = (g.rcu[c] - g.rmu[c] * g.scp) / g.rm;
            const double f2 = (g.rmu[c] - g.rcu[c] * g.scp) / g.rc;
            atomicAdd(&grad[3 * i + c], (f1 + f2) * f0);
            atomicAdd(&grad[3 * (i + 1) + c], f3 * f0);
            atomicAdd(&grad[3 * (i + 2) + c], f3 * f0);
        }
    }
    // block sum of the warps' terms, then one pair of atomics per block
    __shared__ double red[2][32];
    if (lane == 0) { red[0][threadIdx.x >> 5] = e; red[1][threadIdx.x >> 5] = tsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double es = 0.0, ts = 0.0;
        for (int k = 0; k < wpb; k++) { es += red[0][k]; ts += red[1][k]; }
        if (es != 0.0 || ts != 0.0) { atomicAdd(&rst[1], es); atomicAdd(&rst[2 + is], ts); }
        if (blockIdx.x == 0) rst[2 + QNB_MAX_SHELLS + is] = (double)n;
    }
}

}  // namespace qnb
