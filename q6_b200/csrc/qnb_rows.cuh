// qnb_rows.cuh -- round-2 row kernels: FP32 gradient kernels over the chunked unit rows and one FP64 energy kernel
// over flat per-pair lists.
//
// Replaces the bodies of nonbond_pp/_box, nonbond_pw/_box, nonbond_ww/_box, nonbond_ww_spc/_box (nbe, nbe_b, nbe_spc,
// nbe_spcb: nonbondene.f90:4694-5012, 5509-5657, 5886-6077; nonbonded.f90:45-109, 247-291).
//
// Why two kinds of kernel.  The north_star asks for FP32 pair math with FP64 accumulation.  Gradients tolerate FP32
// (bar: 1e-5 relative RMS); the energy sums do not: E%ww%el of a water sphere is ~ -7 kcal/mol out of 5e5 terms of
// magnitude 30-80, so a 1e-6 bar on the net value needs ~1e-10 per term.  Round 1 carried FP64 energy chains inside
// the FP32 kernels: 163-248 registers, 12-16 warps per SM, 63 FP64->FP32 conversions per chunk on the 16-lane XU pipe.
// Here the two jobs are separate kernels that run concurrently and use different pipes:
//   k_water_rows / k_solute_rows  gradient only, pure FP32, packed FFMA2/FMUL2/FADD2 (sm_100) for two site pairs per
//                                 instruction, coordinates as 32-bit fixed point so that i-j differences are exact
//                                 and periodic images come from integer wrap-around; partner records of the next
//                                 chunk loaded into registers while the current one is computed;
//   k_pair_energy                 energies only, pure FP64, one thread per listed pair (owner side), MUFU.RSQ64H seed +
//                                 one Newton step folded into two running sums per charge class.
// The step kernels are written as Body functors (struct XBody { static __device__ void run(args..., BX, NBX, BY, NBY) })
// with a thin __global__ wrapper each, so that the same body can also run inside k_batched (one launch for all windows
// of a batch) and k_uber (several bodies in one launch), qnb_kernels.cuh.
// Each pair's gradient is still evaluated from both sides (full rows, gather only).  Measured alternative
// (tools/microbench2.cu, profiles/r02a_microbench2.txt): returning the partner's nine gradient components through
// warp-coalesced RED.F64 costs 10-16 SM clocks per RED instruction, 90-140 per 32-pair chunk, against ~31 SM clocks to
// recompute the chunk from the other side.
#pragma once
#include "qnb_kernels.cuh"

namespace qnb {

// ---- fixed-point frame: P = frac((x - org) / period) * 2^32 per component.  Differences of two P wrap modulo the
// period, i.e. they ARE the minimum image when period = box length (boxlength*nint(dx/boxlength) of the reference for
// every pair whose components stay clear of half a box length: all listed pairs).  Non-periodic systems use a
// power-of-two period that covers the system several times.
struct FixFrame {
    double org[3], inv_period[3];
    float scale[3];   // period / 2^32
};
__device__ __forceinline__ int fix_coord(double x, double org, double invp) {
    double u = (x - org) * invp;
    u -= floor(u);
    return (int)__double2uint_rd(u * 4294967296.0);   // saturating conversion: u == 1.0 after rounding stays in range
}

constexpr int kRecValid = 0x100;   // rec_i.w = compact type | kRecValid: a zero-filled record is a padding lane of a chunk
typedef float2 f2;
__device__ __forceinline__ f2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 rsq2(f2 a) { return mk2(rsqrt_fast(a.x), rsqrt_fast(a.y)); }

// per packed atom, rewritten every step by k_pack_step
//   rec_i = {P of the atom's unit switch atom (x,y,z), compact type | 0x100}
//   rec_f = {offset from that switch atom (x,y,z), charge}
//   wT    = waters only, three float4 starting at the oxygen's packed index: the oxygen's fixed-point position and the
//           hydrogen offsets t1, t2 from it, laid out as the packed (f2) operands of the row kernels:
//           {Px, Py, t1x, t2x} {Pz, 1, t1y, t2y} {t1z, t2z, 0, 0}        (Px, Py, Pz: int bits)
//   own   = the same three float4 per water, indexed by water number (the own molecule of a water row), the water
//           number in the third one's .z
// and px/py/pz: FP64 coordinates in packed order (energy kernel)
struct PackStepBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(int npk, const FixFrame &F, int nat_solute, const int *__restrict__ pk_atom, const int *__restrict__ pk_sw, const float *__restrict__ pk_q, const int *__restrict__ pk_ct, const double *__restrict__ x, double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz, int4 *__restrict__ rec_i, float4 *__restrict__ rec_f, float4 *__restrict__ wT, float4 *__restrict__ own, double2 *__restrict__ wd, double *__restrict__ zero_out, int nzero, const double *__restrict__ lam_src, double *__restrict__ lam_dst, int nlam, const int BX, const int NBX, const int BY, const int NBY) {
    const int p = BX * blockDim.x + threadIdx.x;
    // batched steps fold the clearing of the output buffer and the upload of lambda into this kernel (two graph nodes less
    // per window); single steps pass null pointers
    if (zero_out) for (int i = p; i < nzero; i += NBX * blockDim.x) zero_out[i] = 0.0;
    if (lam_src && p < nlam) lam_dst[p] = lam_src[p];
    if (p >= npk) return;
    const int i = pk_atom[p], sw = pk_sw[p];
    const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    double xs = xi, ys = yi, zs = zi;
    if (sw != i) { xs = x[3 * sw]; ys = x[3 * sw + 1]; zs = x[3 * sw + 2]; }
    px[p] = xi; py[p] = yi; pz[p] = zi;
    const int Px = fix_coord(xs, F.org[0], F.inv_period[0]), Py = fix_coord(ys, F.org[1], F.inv_period[1]),
              Pz = fix_coord(zs, F.org[2], F.inv_period[2]);
    rec_i[p] = make_int4(Px, Py, Pz, pk_ct[p] | kRecValid);
    rec_f[p] = make_float4((float)(xi - xs), (float)(yi - ys), (float)(zi - zs), pk_q[p]);
    if (i >= nat_solute && sw == i) {
        const float t1x = (float)(x[3 * i + 3] - xi), t1y = (float)(x[3 * i + 4] - yi), t1z = (float)(x[3 * i + 5] - zi);
        const float t2x = (float)(x[3 * i + 6] - xi), t2y = (float)(x[3 * i + 7] - yi), t2z = (float)(x[3 * i + 8] - zi);
        const float4 a = make_float4(__int_as_float(Px), __int_as_float(Py), t1x, t2x);
        const float4 b = make_float4(__int_as_float(Pz), 1.f, t1y, t2y);   // .y: marks a real record (padding lanes read zeros)
        const float4 c = make_float4(t1z, t2z, 0.f, 0.f);
        wT[p] = a; wT[p + 1] = b; wT[p + 2] = c;
        const int w = (i - nat_solute) / 3;
        own[3 * w] = a; own[3 * w + 1] = b; own[3 * w + 2] = make_float4(t1z, t2z, __int_as_float(w), 0.f);
        // FP64 sites of the molecule, contiguous and in packed (cell) order (energy kernel): two double2 slots per
        // packed atom, the water's five {Ox,Oy} {Oz,H1x} {H1y,H1z} {H2x,H2y} {H2z,-} start at its oxygen's
        double2 *o = wd + 2 * (size_t)p;
        o[0] = make_double2(xi, yi); o[1] = make_double2(zi, x[3 * i + 3]); o[2] = make_double2(x[3 * i + 4], x[3 * i + 5]);
        o[3] = make_double2(x[3 * i + 6], x[3 * i + 7]); o[4] = make_double2(x[3 * i + 8], 0.0);
    }
}
};
__global__ void __launch_bounds__(256)
k_pack_step(int npk, FixFrame F, int nat_solute, const int *__restrict__ pk_atom, const int *__restrict__ pk_sw,
            const float *__restrict__ pk_q, const int *__restrict__ pk_ct, const double *__restrict__ x,
            double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz, int4 *__restrict__ rec_i,
            float4 *__restrict__ rec_f, float4 *__restrict__ wT, float4 *__restrict__ own, double2 *__restrict__ wd,
            double *__restrict__ zero_out, int nzero, const double *__restrict__ lam_src, double *__restrict__ lam_dst, int nlam) {
    PackStepBody::run(npk, F, nat_solute, pk_atom, pk_sw, pk_q, pk_ct, x, px, py, pz, rec_i, rec_f, wT, own, wd, zero_out, nzero, lam_src, lam_dst, nlam, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// FP32 parameters of the row kernels (kernel argument, constant bank)
struct RowPar {
    float scale[3];
    // water-water, packed by the groups of the kernel: g=0 (a in {1,2}, b=0), 1 (1,1)(2,2), 2 (1,2)(2,1), 3 (0,1)(0,2)
    f2 wQ[4], wA12[4], wB6[4];
    float q00, A00, B00;          // (0,0): charge product, 12 A, 6 B
    float wq0;                    // charges of the three water sites (pw)
    f2 wq12;
    float el14;
    int nct;
};

// LJ + Coulomb scale factor of one site pair: returns c with grad_i += d * c  (c = -dv of nbe, nonbonded.f90:45-75)
__device__ __forceinline__ float pair_c(float r2, float qq, float A12, float B6) {
    const float ri = rsqrt_fast(r2), ri2 = ri * ri, r6 = ri2 * ri2 * ri2;
    const float w = fmaf(A12 * r6, r6, fmaf(-B6, r6, qq * ri));
    return w * ri2;
}
__device__ __forceinline__ f2 pair_c2(f2 r2, f2 qq, f2 A12, f2 B6) {
    const f2 ri = rsq2(r2), ri2 = mul2(ri, ri), r6 = mul2(mul2(ri2, ri2), ri2);
    const f2 w = fma2(mul2(A12, r6), r6, fma2(mk2(-B6.x, -B6.y), r6, mul2(qq, ri)));
    return mul2(w, ri2);
}
__device__ __forceinline__ f2 coul_c2(f2 r2, f2 qq) {
    const f2 ri = rsq2(r2), ri2 = mul2(ri, ri);
    return mul2(mul2(qq, ri), ri2);
}
__device__ __forceinline__ f2 len2(f2 dx, f2 dy, f2 dz) { return fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))); }

// ------------------------------------------------------------------------------------------------
// Water rows, gradient only: ww (3x3 site tile, both sides of every pair) + the water side of pw.
// One warp streams a contiguous range of 32-entry chunks (k_warp_starts); lanes = partners of the row's water.
//
// The partners' records are gathers (cell-ordered, so neighbouring lanes hit neighbouring records, but still one L1/L2
// round trip per chunk): r02c measured 60 % of the stall samples on their first consumers.  What was tried:
//  * r02e-r02g: every lane copies its partner's record into its own shared-memory slot with cp.async, three chunks deep.
//    Latency stalls fell (long scoreboard 7.9 -> 2.6 per issue) but the time only went 63 -> 55 us whatever the occupancy
//    (4, 5, 6 blocks per SM): a GATHERED cp.async writes its sectors to shared memory as they arrive, ~11 data-pipe
//    wavefronts per instruction, and l1tex__data_pipe_lsu_wavefronts sat at 81 % of peak;
//  * now: the records of chunk c+1 are loaded into registers (three LDG.128) while chunk c is computed -- two register
//    sets that swap roles in a loop unrolled twice -- entries three chunks ahead, and cp.async only where it is cheap:
//    the own molecule's record of the next row (lanes 0-2, a ring of four slots).
constexpr uint32_t kKindBit = kSpecialBit;   // water-row chunk entries: set in every lane of a B chunk (k_chunk_fill)
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void *gsrc, unsigned nbytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// the chunk entries stream from DRAM (21 MB per step on the 98 k-atom box): their lines are pulled into L2 kWPrefetch
// chunks ahead, so that the register prefetch only has to cover an L2 hit
constexpr int kWPrefetch = 10;
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
struct WRec { float4 a, b, c; };   // a partner's record: water {wT[p], wT[p+1], wT[p+2]}, solute atom {rec_i[p], rec_f[p], -}

// MINB: resident blocks per SM the register allocation aims at (QNB_WROWS_MINB selects at run time)
template <bool SPC, int MINB>
struct WaterRowsBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(const RowPar &P, const int4 *__restrict__ rec_i, const float4 *__restrict__ rec_f, const float4 *__restrict__ wT, const float4 *__restrict__ own, const float2 *__restrict__ pw0, const float4 *__restrict__ pw12, const int *__restrict__ wstart, const int2 *__restrict__ cdesc, const uint32_t *__restrict__ crow, int i0_water, double *__restrict__ grad, const int BX, const int NBX, const int BY, const int NBY) {
    __shared__ float4 Own[4][4][4];            // [warp][slot][record part]; the issue side runs at most three row changes ahead
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = (BX * blockDim.x + threadIdx.x) >> 5;
    const int c0 = wstart[gw], c1 = wstart[gw + 1];
    if (c0 >= c1) return;
    const unsigned sOwn = smem_addr(&Own[wib][0][0]);
    int cur_w = -1;
    int Pix = 0, Piy = 0, Piz = 0;
    f2 nSx = mk2(0, 0), nSy = nSx, nSz = nSx;                 // minus the own hydrogens' offsets {s1, s2}
    f2 nWx = nSx, nWy = nSx, nWz = nSx;                       // the same, swapped: {s2, s1}
    // gradient of the own sites {1,2}, {2,1} (from the swapped operands) and of site 0 (two halves + the O-O term)
    f2 G12x = nSx, G12y = nSx, G12z = nSx, G21x = nSx, G21y = nSx, G21z = nSx, G0x = nSx, G0y = nSx, G0z = nSx;
    float g0x = 0.f, g0y = 0.f, g0z = 0.f;
    const int slot = split_slot<9, 16>(lane);
    const f2 wQ2 = mk2(P.wQ[2].y, P.wQ[2].x), wA2 = mk2(P.wA12[2].y, P.wA12[2].x), wB2 = mk2(P.wB6[2].y, P.wB6[2].x);

    auto flush = [&]() {
        if (cur_w < 0) return;
        float s[9] = {g0x + (G0x.x + G0x.y), g0y + (G0y.x + G0y.y), g0z + (G0z.x + G0z.y),
                      G12x.x + G21x.y, G12y.x + G21y.y, G12z.x + G21z.y, G12x.y + G21x.x, G12y.y + G21y.x, G12z.y + G21z.x};
        const float mine = split_reduce<9, 16>(s, lane);
        if (slot >= 0) atomicAdd(&grad[3 * (size_t)(i0_water + 3 * cur_w) + slot], (double)mine);
    };
    const int *__restrict__ cunit = reinterpret_cast<const int *>(cdesc);
    // Pipeline state.  e1: entry of chunk c+1 (its records are loaded while chunk c is computed); e2, e3: entries of
    // c+2, c+3 in flight; u1, u2, u3: units of c+1, c+2, c+3.  kinds/rows: one bit per chunk ahead (bit k = chunk c+k):
    // B chunk / first chunk of a row.
    uint32_t e1, e2, e3 = 0xffffffffu;
    int u1, u2, u3 = -1;
    unsigned kinds = 0, rows = 0, islot = 0, cslot = 0;
    int ie = (c0 + 3) * 32 + lane;   // entry index of chunk c+3
    // records of the partner named by entry e; a padding lane names the all-zero record behind the last packed atom
    // (b.y / a.w mark real records), so the loads need no test
    auto load_rec = [&](uint32_t e, unsigned bit) -> WRec {
        WRec r;
        const int p = (int)(e & kIdMask);
        if (!__any_sync(kFull, (e & kKindBit) != 0)) {
            const float4 *src = wT + p;
            r.a = src[0]; r.b = src[1]; r.c = src[2];
        } else {
            const int4 ri = rec_i[p];
            r.a = make_float4(__int_as_float(ri.x), __int_as_float(ri.y), __int_as_float(ri.z), __int_as_float(ri.w));
            r.b = rec_f[p];
            r.c = r.b;
            kinds |= bit;
        }
        return r;
    };
    // own record of the row that starts with chunk cc (unit u, previous chunk's unit uprev): lanes 0-2, asynchronous
    auto own_prefetch = [&](int cc, int u, int uprev, unsigned bit) {
        if (cc < c1 && u != uprev) {
            if (lane < 3) cp_async16(sOwn + islot * 64 + lane * 16, own + 3 * (size_t)u + lane, 16u);
            islot = (islot + 1) & 3;
            rows |= bit;
        }
        cp_async_commit();
    };
    auto compute = [&](const WRec &R) {
        cp_async_wait<2>();
        if (rows & 1u) {
            flush();
            __syncwarp();   // lanes 0-2 have passed their wait: the own record is complete
            const unsigned so = sOwn + cslot * 64;
            const float4 a = lds128(so), b = lds128(so + 16), cc = lds128(so + 32);
            cslot = (cslot + 1) & 3;
            cur_w = __float_as_int(cc.z);
            Pix = __float_as_int(a.x); Piy = __float_as_int(a.y); Piz = __float_as_int(b.x);
            nSx = mk2(-a.z, -a.w); nSy = mk2(-b.z, -b.w); nSz = mk2(-cc.x, -cc.y);
            nWx = mk2(-a.w, -a.z); nWy = mk2(-b.w, -b.z); nWz = mk2(-cc.y, -cc.x);
            G12x = G12y = G12z = G21x = G21y = G21z = G0x = G0y = G0z = mk2(0.f, 0.f);
            g0x = g0y = g0z = 0.f;
        }
        const float4 r0 = R.a, r1 = R.b;
        if (!(kinds & 1u)) {
            const float4 r2 = R.c;
            // vector from the own oxygen to the partner's; a padding lane (zero record) is sent to 1e18 A, where
            // every r^-3 and r^-6 below flushes to zero
            float Rx = (float)(__float_as_int(r0.x) - Pix) * P.scale[0];
            const float Ry = (float)(__float_as_int(r0.y) - Piy) * P.scale[1], Rz = (float)(__float_as_int(r1.x) - Piz) * P.scale[2];
            Rx = r1.y != 0.f ? Rx : 1.0e18f;
            const f2 RRx = mk2(Rx, Rx), RRy = mk2(Ry, Ry), RRz = mk2(Rz, Rz);
            // partner hydrogens {1, 2} seen from the own oxygen
            const f2 Ux = add2(RRx, mk2(r0.z, r0.w)), Uy = add2(RRy, mk2(r1.z, r1.w)), Uz = add2(RRz, mk2(r2.x, r2.y));
            f2 dx, dy, dz, cc;
            // own {1,2} - partner 0
            dx = add2(RRx, nSx); dy = add2(RRy, nSy); dz = add2(RRz, nSz);
            cc = SPC ? coul_c2(len2(dx, dy, dz), P.wQ[0]) : pair_c2(len2(dx, dy, dz), P.wQ[0], P.wA12[0], P.wB6[0]);
            G12x = fma2(dx, cc, G12x); G12y = fma2(dy, cc, G12y); G12z = fma2(dz, cc, G12z);
            // own {1,2} - partner {1,2}
            dx = add2(Ux, nSx); dy = add2(Uy, nSy); dz = add2(Uz, nSz);
            cc = SPC ? coul_c2(len2(dx, dy, dz), P.wQ[1]) : pair_c2(len2(dx, dy, dz), P.wQ[1], P.wA12[1], P.wB6[1]);
            G12x = fma2(dx, cc, G12x); G12y = fma2(dy, cc, G12y); G12z = fma2(dz, cc, G12z);
            // own {2,1} - partner {1,2}: the own offsets are swapped (once per row), not the partner's (once per chunk)
            dx = add2(Ux, nWx); dy = add2(Uy, nWy); dz = add2(Uz, nWz);
            cc = SPC ? coul_c2(len2(dx, dy, dz), wQ2) : pair_c2(len2(dx, dy, dz), wQ2, wA2, wB2);
            G21x = fma2(dx, cc, G21x); G21y = fma2(dy, cc, G21y); G21z = fma2(dz, cc, G21z);
            // own 0 - partner {1,2}
            cc = SPC ? coul_c2(len2(Ux, Uy, Uz), P.wQ[3]) : pair_c2(len2(Ux, Uy, Uz), P.wQ[3], P.wA12[3], P.wB6[3]);
            G0x = fma2(Ux, cc, G0x); G0y = fma2(Uy, cc, G0y); G0z = fma2(Uz, cc, G0z);
            // (0,0): the pair that carries LJ in nonbond_ww_spc
            const float c00 = pair_c(fmaf(Rx, Rx, fmaf(Ry, Ry, Rz * Rz)), P.q00, P.A00, P.B00);
            g0x = fmaf(Rx, c00, g0x); g0y = fmaf(Ry, c00, g0y); g0z = fmaf(Rz, c00, g0z);
        } else {
            // solute atom acting on the own water (pw seen from the water: gradient only)
            const int ctw = __float_as_int(r0.w);
            const float4 oj = r1;
            float Rx = (float)(__float_as_int(r0.x) - Pix) * P.scale[0];
            const float Ry = (float)(__float_as_int(r0.y) - Piy) * P.scale[1], Rz = (float)(__float_as_int(r0.z) - Piz) * P.scale[2];
            Rx = (ctw & kRecValid) ? Rx : 1.0e18f;
            const float ex = Rx + oj.x, ey = Ry + oj.y, ez = Rz + oj.z;
            const float2 l0 = pw0[ctw & 0xff];
            const float4 l12 = pw12[ctw & 0xff];
            const f2 dx = add2(mk2(ex, ex), nSx), dy = add2(mk2(ey, ey), nSy), dz = add2(mk2(ez, ez), nSz);
            const f2 cc = pair_c2(len2(dx, dy, dz), mul2(P.wq12, mk2(oj.w, oj.w)), mk2(l12.x, l12.y), mk2(l12.z, l12.w));
            G12x = fma2(dx, cc, G12x); G12y = fma2(dy, cc, G12y); G12z = fma2(dz, cc, G12z);
            const float c0q = pair_c(fmaf(ex, ex, fmaf(ey, ey, ez * ez)), P.wq0 * oj.w, l0.x, l0.y);
            g0x = fmaf(ex, c0q, g0x); g0y = fmaf(ey, c0q, g0y); g0z = fmaf(ez, c0q, g0z);
        }
    };
    WRec RA, RB;
    {
        const uint32_t e0 = crow[(size_t)c0 * 32 + lane];
        e1 = c0 + 1 < c1 ? crow[(size_t)(c0 + 1) * 32 + lane] : 0xffffffffu;
        e2 = c0 + 2 < c1 ? crow[(size_t)(c0 + 2) * 32 + lane] : 0xffffffffu;
        const int u0 = cunit[2 * (size_t)c0];
        u1 = c0 + 1 < c1 ? cunit[2 * (size_t)(c0 + 1)] : -1;
        u2 = c0 + 2 < c1 ? cunit[2 * (size_t)(c0 + 2)] : -1;
        own_prefetch(c0, u0, -1, 1u);
        own_prefetch(c0 + 1, u1, u0, 2u);
        RA = load_rec(e0, 1u);
    }
    int c = c0;
    // one step: records of chunk c+1 into NXT, entry/unit of chunk c+3, own record of chunk c+2's row, compute chunk c
#define QNB_WSTEP(CUR, NXT, PF)                                                                           \
    {                                                                                                   \
        if (c + 1 < c1) NXT = load_rec(e1, 2u);                                                         \
        if (c + 3 < c1) { e3 = crow[ie]; u3 = cunit[2 * (size_t)(c + 3)]; }                             \
        if (PF && lane == 0 && c + kWPrefetch < c1) { prefetch_l2(crow + (size_t)(c + kWPrefetch) * 32); prefetch_l2(crow + (size_t)(c + kWPrefetch) * 32 + 32); } \
        own_prefetch(c + 2, u2, u1, 4u);                                                                \
        compute(CUR);                                                                                   \
        kinds >>= 1; rows >>= 1;                                                                        \
        e1 = e2; e2 = e3; u1 = u2; u2 = u3; ie += 32;                                                   \
        if (++c >= c1) break;                                                                           \
    }
    for (;;) {
        QNB_WSTEP(RA, RB, true)
        QNB_WSTEP(RB, RA, false)
    }
#undef QNB_WSTEP
    flush();
}
};
template <bool SPC, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_water_rows(RowPar P, const int4 *__restrict__ rec_i, const float4 *__restrict__ rec_f, const float4 *__restrict__ wT,
             const float4 *__restrict__ own, const float2 *__restrict__ pw0, const float4 *__restrict__ pw12,
             const int *__restrict__ wstart, const int2 *__restrict__ cdesc, const uint32_t *__restrict__ crow,
             int i0_water /* nat_solute */, double *__restrict__ grad) {
    WaterRowsBody<SPC, MINB>::run(P, rec_i, rec_f, wT, own, pw0, pw12, wstart, cdesc, crow, i0_water, grad, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// ------------------------------------------------------------------------------------------------
// Solute rows, gradient only: pp (both sides) + the solute side of pw.  A chunk belongs to one (charge group, tile of
// up to four non-Q atoms); the tile's atoms are the packed halves: {t0,t1} and {t2,t3}.
// ljp: [2][nct][nct] {12 A, 6 B} of a type pair, plane 0 by the pair's own LJ code (ljcod), plane 1 for 1-4 pairs
// (code 3); both combination rules are resolved on the host.
#ifndef QNB_SROWS_MINB
#define QNB_SROWS_MINB 4
#endif
template <bool HLJ /* solvent hydrogens carry LJ against some solute type */>
struct SoluteRowsBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(const RowPar &P, const int *__restrict__ upk, const int *__restrict__ nq_off, const int4 *__restrict__ rec_i, const float4 *__restrict__ rec_f, const float4 *__restrict__ wT, const float2 *__restrict__ ljp, const float2 *__restrict__ pw0, const float4 *__restrict__ pw12, const int *__restrict__ wstart, const int2 *__restrict__ cdesc, const uint32_t *__restrict__ crow, const uint16_t *__restrict__ cspec, const int *__restrict__ pk_atom, double *__restrict__ grad, const int BX, const int NBX, const int BY, const int NBY) {
    const int lane = threadIdx.x & 31;
    const int gw = (BX * blockDim.x + threadIdx.x) >> 5;
    const int c0 = wstart[gw], c1 = wstart[gw + 1];
    if (c0 >= c1) return;
    int cur_key = -1, nt = 0, pi0 = 0;
    int Pix = 0, Piy = 0, Piz = 0;
    f2 nOx[2], nOy[2], nOz[2], Q[2];      // minus the tile atoms' offsets from the switch atom; charges
    int ct[4] = {0, 0, 0, 0};
    f2 Gx[2] = {mk2(0.f, 0.f), mk2(0.f, 0.f)}, Gy[2] = {mk2(0.f, 0.f), mk2(0.f, 0.f)}, Gz[2] = {mk2(0.f, 0.f), mk2(0.f, 0.f)};
    const int slot = split_slot<12, 16>(lane);
    auto flush = [&]() {
        if (cur_key < 0) return;
        float s[12] = {Gx[0].x, Gy[0].x, Gz[0].x, Gx[0].y, Gy[0].y, Gz[0].y, Gx[1].x, Gy[1].x, Gz[1].x, Gx[1].y, Gy[1].y, Gz[1].y};
        const float mine = split_reduce<12, 16>(s, lane);
        const int t = slot / 3;
        if (slot >= 0 && t < nt) atomicAdd(&grad[3 * (size_t)pk_atom[pi0 + t] + (slot - 3 * t)], (double)mine);
    };
    int2 dn = cdesc[c0];
    uint32_t en = crow[(size_t)c0 * 32 + lane];
    uint32_t sn = cspec[(size_t)c0 * 32 + lane];
    for (int c = c0; c < c1; c++) {
        const int2 d = dn;
        const uint32_t e = en, spec = sn;
        if (c + 1 < c1) { dn = cdesc[c + 1]; en = crow[(size_t)(c + 1) * 32 + lane]; sn = cspec[(size_t)(c + 1) * 32 + lane]; }
        const int tile = (d.y >> 8) & 0xff;
        const int key = d.x * 256 + tile;
        if (key != cur_key) {
            flush();
            cur_key = key;
            const int k0 = nq_off[d.x] + tile * 4;
            nt = min(4, nq_off[d.x + 1] - k0);
            pi0 = upk[d.x] + tile * 4;
            const int4 r0 = rec_i[pi0];
            Pix = r0.x; Piy = r0.y; Piz = r0.z;
            float ox[4], oy[4], oz[4], q[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                // atoms beyond the tile repeat the last one with zero charge and a type whose row is never kept
                const int pt = pi0 + min(t, nt - 1);
                const float4 o = rec_f[pt];
                ox[t] = -o.x; oy[t] = -o.y; oz[t] = -o.z; q[t] = t < nt ? o.w : 0.f;
                ct[t] = rec_i[pt].w & 0xff;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                nOx[h] = mk2(ox[2 * h], ox[2 * h + 1]); nOy[h] = mk2(oy[2 * h], oy[2 * h + 1]); nOz[h] = mk2(oz[2 * h], oz[2 * h + 1]);
                Q[h] = mk2(q[2 * h], q[2 * h + 1]);
                Gx[h] = Gy[h] = Gz[h] = mk2(0.f, 0.f);
            }
        }
        const bool valid = e != kPadEntry;
        const int p = valid ? (int)(e & kIdMask) : 0;
        const int4 rj = rec_i[p];
        float Rx = (float)(rj.x - Pix) * P.scale[0];
        const float Ry = (float)(rj.y - Piy) * P.scale[1], Rz = (float)(rj.z - Piz) * P.scale[2];
        Rx = valid ? Rx : 1.0e18f;
        if ((d.y & 0xff) == kChunkA) {
            // ---- solute partner atom: 3 bits per tile atom from k_chunk_fill: 1 skip, 2 1-4 pair, (4 energy side)
            const float4 oj = rec_f[p];
            const float ex = Rx + oj.x, ey = Ry + oj.y, ez = Rz + oj.z;
            const f2 EEx = mk2(ex, ex), EEy = mk2(ey, ey), EEz = mk2(ez, ez);
            const int plane = P.nct * P.nct;
            float A12[4], B6[4], qs[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const unsigned sb = (spec >> (3 * t)) & 7u;
                const bool i14 = (sb & 2u) != 0, skip = (sb & 1u) != 0 || t >= nt;
                const float2 l = ljp[(i14 ? plane : 0) + ct[t] * P.nct + (rj.w & 0xff)];
                A12[t] = skip ? 0.f : l.x; B6[t] = skip ? 0.f : l.y;
                qs[t] = skip ? 0.f : (i14 ? oj.w * P.el14 : oj.w);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const f2 dx = add2(EEx, nOx[h]), dy = add2(EEy, nOy[h]), dz = add2(EEz, nOz[h]);
                // r = 0 only for an atom against itself (skip): keep its 1/r finite
                f2 r2 = len2(dx, dy, dz);
                r2 = mk2(fmaxf(r2.x, 1e-12f), fmaxf(r2.y, 1e-12f));
                const f2 cc = pair_c2(r2, mul2(Q[h], mk2(qs[2 * h], qs[2 * h + 1])), mk2(A12[2 * h], A12[2 * h + 1]), mk2(B6[2 * h], B6[2 * h + 1]));
                Gx[h] = fma2(dx, cc, Gx[h]); Gy[h] = fma2(dy, cc, Gy[h]); Gz[h] = fma2(dz, cc, Gz[h]);
            }
        } else {
            // ---- partner water: the three sites against every tile atom (nonbond_pw, solute side)
            const float4 ta = wT[p], tb = wT[p + 1], tc = wT[p + 2];
            const float ux[3] = {Rx, Rx + ta.z, Rx + ta.w}, uy[3] = {Ry, Ry + tb.z, Ry + tb.w}, uz[3] = {Rz, Rz + tc.x, Rz + tc.y};
            float2 l0[4];
            float4 l12[4];
#pragma unroll
            for (int t = 0; t < 4; t++) { l0[t] = pw0[ct[t]]; if (HLJ) l12[t] = pw12[ct[t]]; }
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const f2 dx = add2(mk2(ux[b], ux[b]), nOx[h]), dy = add2(mk2(uy[b], uy[b]), nOy[h]), dz = add2(mk2(uz[b], uz[b]), nOz[h]);
                    const float qb = b == 0 ? P.wq0 : (b == 1 ? P.wq12.x : P.wq12.y);
                    const f2 qq = mul2(Q[h], mk2(qb, qb));
                    f2 cc;
                    if (b == 0) cc = pair_c2(len2(dx, dy, dz), qq, mk2(l0[2 * h].x, l0[2 * h + 1].x), mk2(l0[2 * h].y, l0[2 * h + 1].y));
                    else if (HLJ) {
                        const f2 A = b == 1 ? mk2(l12[2 * h].x, l12[2 * h + 1].x) : mk2(l12[2 * h].y, l12[2 * h + 1].y);
                        const f2 B = b == 1 ? mk2(l12[2 * h].z, l12[2 * h + 1].z) : mk2(l12[2 * h].w, l12[2 * h + 1].w);
                        cc = pair_c2(len2(dx, dy, dz), qq, A, B);
                    } else cc = coul_c2(len2(dx, dy, dz), qq);
                    Gx[h] = fma2(dx, cc, Gx[h]); Gy[h] = fma2(dy, cc, Gy[h]); Gz[h] = fma2(dz, cc, Gz[h]);
                }
            }
        }
    }
    flush();
}
};
template <bool HLJ /* solvent hydrogens carry LJ against some solute type */>
__global__ void __launch_bounds__(128, QNB_SROWS_MINB)
k_solute_rows(RowPar P, const int *__restrict__ upk, const int *__restrict__ nq_off, const int4 *__restrict__ rec_i,
              const float4 *__restrict__ rec_f, const float4 *__restrict__ wT, const float2 *__restrict__ ljp,
              const float2 *__restrict__ pw0, const float4 *__restrict__ pw12, const int *__restrict__ wstart,
              const int2 *__restrict__ cdesc, const uint32_t *__restrict__ crow, const uint16_t *__restrict__ cspec,
              const int *__restrict__ pk_atom, double *__restrict__ grad) {
    SoluteRowsBody<HLJ>::run(P, upk, nq_off, rec_i, rec_f, wT, ljp, pw0, pw12, wstart, cdesc, crow, cspec, pk_atom, grad, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// ------------------------------------------------------------------------------------------------
// Energies: one thread per listed pair (the reference's own list: each pair once), FP64.
//   ww_pairs {pi, pj}: packed indices of the two oxygens                      (E%ww, nonbond_ww(_spc)(_box))
//   pp_pairs {pi, pj | img << 24 | 1-4 << 30 | skip << 31}: packed atoms      (E%pp, nonbond_pp(_box))
//   pw_pairs {pi, pj | img << 24}: packed solute atom, packed water oxygen    (E%pw, nonbond_pw(_box))
// 1/r: MUFU.RSQ64H seed y0 (relative error e ~ 2^-21) and one Newton step, y = y0 (3/2 - r2 y0^2 / 2), error 3/2 e^2.
struct EnergyPar {
    double wwQ[9], wwA[9], wwB[9];   // water-water site pairs (a*3+b)
    double wq[3];
    double el14;
    int wct[3];
    int nct, geometric, any_atom, spc;
    double box[3], inv_box[3];
};
__device__ __forceinline__ double rsq64_seed(double r2) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    return y;
}
__device__ __forceinline__ double rsq64(double r2) {
    const double y0 = rsq64_seed(r2);
    const double p = r2 * y0;
    return y0 * fma(-0.5 * p, y0, 1.5);
}
__device__ __forceinline__ void lj_f64(const EnergyPar &P, const double *__restrict__ ljd, int cta, int ctb, int code, double &A, double &B) {
    const double ax = ljd[(cta * 3 + code - 1) * 2], ay = ljd[(cta * 3 + code - 1) * 2 + 1];
    const double bx = ljd[(ctb * 3 + code - 1) * 2], by = ljd[(ctb * 3 + code - 1) * 2 + 1];
    if (P.geometric) { A = ax * bx; B = ay * by; }
    else { double t = ax + bx; t = t * t; t = t * t * t; const double e = ay * by; A = t * t * e; B = 2.0 * t * e; }
}

template <bool PBC>
struct PairEnergyBody {
    // BX, NBX, BY, NBY: block index and grid size (the batched launch passes those of the window, k_batched)
    static __device__ __forceinline__ void run(const EnergyPar &P, int n_ww, const int2 *__restrict__ ww_pairs, int n_pp, const int2 *__restrict__ pp_pairs, int n_pw, const int2 *__restrict__ pw_pairs, const double *__restrict__ px, const double *__restrict__ py, const double *__restrict__ pz, const double *__restrict__ pk_qd, const int *__restrict__ pk_ct, const int *__restrict__ pk_sw, const double *__restrict__ x, const double2 *__restrict__ wd, const double *__restrict__ ljd, const uint8_t *__restrict__ ljcode, double *__restrict__ Eslots, int nE, const int BX, const int NBX, const int BY, const int NBY) {
    const int tid = BX * blockDim.x + threadIdx.x, nthr = NBX * blockDim.x;
    double e_ww_el = 0.0, e_ww_vdw = 0.0, e_pp_el = 0.0, e_pp_vdw = 0.0, e_pw_el = 0.0, e_pw_vdw = 0.0;
    // ---- water-water
    {
        // SPC-type water with equal hydrogens: sum 1/r per charge class (OO, OH, HH) and weigh at the end
        const bool classes = P.spc && P.wwQ[1] == P.wwQ[2] && P.wwQ[1] == P.wwQ[3] && P.wwQ[1] == P.wwQ[6] && P.wwQ[4] == P.wwQ[5] &&
                             P.wwQ[4] == P.wwQ[7] && P.wwQ[4] == P.wwQ[8];
        double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
        for (int t = tid; t < n_ww; t += nthr) {
            const int2 pr = ww_pairs[t];
            double xi[3][3], xj[3][3];
            {
                // (staging the partner's sites through shared memory with cp.async was measured slower, r02h: 64 vs 43 us
                // on the 98 k-atom box -- a gathered cp.async costs ~11 L1 data-pipe wavefronts per instruction)
                const double2 *wi = wd + 2 * (size_t)pr.x, *wj = wd + 2 * (size_t)pr.y;
                const double2 a0 = wi[0], a1 = wi[1], a2 = wi[2], a3 = wi[3], a4 = wi[4];
                const double2 b0 = wj[0], b1 = wj[1], b2 = wj[2], b3 = wj[3], b4 = wj[4];
                xi[0][0] = a0.x; xi[0][1] = a0.y; xi[0][2] = a1.x; xi[1][0] = a1.y; xi[1][1] = a2.x; xi[1][2] = a2.y;
                xi[2][0] = a3.x; xi[2][1] = a3.y; xi[2][2] = a4.x;
                xj[0][0] = b0.x; xj[0][1] = b0.y; xj[0][2] = b1.x; xj[1][0] = b1.y; xj[1][1] = b2.x; xj[1][2] = b2.y;
                xj[2][0] = b3.x; xj[2][1] = b3.y; xj[2][2] = b4.x;
            }
            if (PBC) {
                // one shift per molecule pair from the O-O vector (nonbond_ww_spc_box L6016-6017, nonbond_ww_box); rint
                // instead of nint: they differ only for a pair exactly half a box apart, which no list holds
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double sh = P.box[k] * rint((xi[0][k] - xj[0][k]) * P.inv_box[k]);
                    xj[0][k] += sh; xj[1][k] += sh; xj[2][k] += sh;
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) {
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const double dx = xj[b][0] - xi[a][0], dy = xj[b][1] - xi[a][1], dz = xj[b][2] - xi[a][2];
                    const double r2 = fma(dx, dx, fma(dy, dy, dz * dz));
                    const bool lj = !P.spc || (a == 0 && b == 0);
                    if (classes && !lj) {
                        const int cl = (a == 0 || b == 0) ? 1 : 2;
                        const double y0 = rsq64_seed(r2);
                        s1[cl] += y0;
                        s2[cl] = fma(y0 * y0, r2 * y0, s2[cl]);
                    } else {
                        const double y = rsq64(r2);
                        e_ww_el = fma(P.wwQ[a * 3 + b], y, e_ww_el);
                        if (lj) {
                            const double y2 = y * y, r6 = y2 * y2 * y2;
                            e_ww_vdw += fma(P.wwA[a * 3 + b] * r6, r6, -P.wwB[a * 3 + b] * r6);
                        }
                    }
                }
            }
        }
        e_ww_el += P.wwQ[1] * (1.5 * s1[1] - 0.5 * s2[1]) + P.wwQ[4] * (1.5 * s1[2] - 0.5 * s2[2]);
    }
    // ---- solute-solute
    for (int t = tid; t < n_pp; t += nthr) {
        const int2 pr = pp_pairs[t];
        if (pr.y < 0) continue;   // excluded pair, or the other orientation of a pair inside one group
        const int pj = pr.y & (int)kIdMask;
        const bool i14 = (pr.y >> 30) & 1;
        double dx = px[pj] - px[pr.x], dy = py[pj] - py[pr.x], dz = pz[pj] - pz[pr.x];
        if (PBC) {
            if (P.any_atom) {
                dx += P.box[0] * img_comp((uint32_t)pr.y, 0); dy += P.box[1] * img_comp((uint32_t)pr.y, 1); dz += P.box[2] * img_comp((uint32_t)pr.y, 2);
            } else {
                // nonbond_pp_box L4791-4801: shift = boxlength*nint((x(sw_i)-x(sw_j))*inv_boxl)
                const int si = pk_sw[pr.x], sj = pk_sw[pj];
                dx += pshift(x[3 * si] - x[3 * sj], P.box[0], P.inv_box[0]);
                dy += pshift(x[3 * si + 1] - x[3 * sj + 1], P.box[1], P.inv_box[1]);
                dz += pshift(x[3 * si + 2] - x[3 * sj + 2], P.box[2], P.inv_box[2]);
            }
        }
        const int cti = pk_ct[pr.x], ctj = pk_ct[pj];
        double A, B;
        lj_f64(P, ljd, cti, ctj, i14 ? 3 : (int)ljcode[cti * P.nct + ctj], A, B);
        const double qq = i14 ? pk_qd[pr.x] * pk_qd[pj] * P.el14 : pk_qd[pr.x] * pk_qd[pj];
        const double y = rsq64(fma(dx, dx, fma(dy, dy, dz * dz)));
        const double y2 = y * y, r6 = y2 * y2 * y2;
        e_pp_el = fma(qq, y, e_pp_el);
        e_pp_vdw += fma(A * r6, r6, -B * r6);
    }
    // ---- solute-water
    for (int t = tid; t < n_pw; t += nthr) {
        const int2 pr = pw_pairs[t];
        const int pj = pr.y & (int)kIdMask;
        double sh[3] = {0, 0, 0};
        const double xi = px[pr.x], yi = py[pr.x], zi = pz[pr.x];
        if (PBC) {
            if (P.any_atom) {
                sh[0] = P.box[0] * img_comp((uint32_t)pr.y, 0); sh[1] = P.box[1] * img_comp((uint32_t)pr.y, 1); sh[2] = P.box[2] * img_comp((uint32_t)pr.y, 2);
            } else {
                // nonbond_pw_box: shift = boxlength*nint((x(solute switch)-x(water O))*inv_boxl)
                const int si = pk_sw[pr.x];
                sh[0] = pshift(x[3 * si] - px[pj], P.box[0], P.inv_box[0]);
                sh[1] = pshift(x[3 * si + 1] - py[pj], P.box[1], P.inv_box[1]);
                sh[2] = pshift(x[3 * si + 2] - pz[pj], P.box[2], P.inv_box[2]);
            }
        }
        const int cti = pk_ct[pr.x];
        const double qi = pk_qd[pr.x];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const double dx = (px[pj + b] - xi) + sh[0], dy = (py[pj + b] - yi) + sh[1], dz = (pz[pj + b] - zi) + sh[2];
            double A, B;
            lj_f64(P, ljd, cti, P.wct[b], (int)ljcode[cti * P.nct + P.wct[b]], A, B);
            const double y = rsq64(fma(dx, dx, fma(dy, dy, dz * dz)));
            const double y2 = y * y, r6 = y2 * y2 * y2;
            e_pw_el = fma(qi * P.wq[b], y, e_pw_el);
            e_pw_vdw += fma(A * r6, r6, -B * r6);
        }
    }
    // ---- block sums, one set of atomics per block
    __shared__ double red[4][6];
    double v[6] = {e_pp_el, e_pp_vdw, e_pw_el, e_pw_vdw, e_ww_el, e_ww_vdw};
#pragma unroll
    for (int k = 0; k < 6; k++) v[k] = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 6; k++) red[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 6) {
        const double s = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
        if (s != 0.0) atomicAdd(&Eslots[(size_t)(BX & (kESlots - 1)) * nE + threadIdx.x], s);
    }
}
};
template <bool PBC>
__global__ void __launch_bounds__(128)
k_pair_energy(EnergyPar P, int n_ww, const int2 *__restrict__ ww_pairs, int n_pp, const int2 *__restrict__ pp_pairs,
              int n_pw, const int2 *__restrict__ pw_pairs, const double *__restrict__ px, const double *__restrict__ py,
              const double *__restrict__ pz, const double *__restrict__ pk_qd, const int *__restrict__ pk_ct,
              const int *__restrict__ pk_sw, const double *__restrict__ x, const double2 *__restrict__ wd,
              const double *__restrict__ ljd, const uint8_t *__restrict__ ljcode, double *__restrict__ Eslots, int nE) {
    PairEnergyBody<PBC>::run(P, n_ww, ww_pairs, n_pp, pp_pairs, n_pw, pw_pairs, px, py, pz, pk_qd, pk_ct, pk_sw, x, wd, ljd, ljcode, Eslots, nE, blockIdx.x, gridDim.x, blockIdx.y, gridDim.y);
}

// ---- flat pair lists of the energy kernel, written at list-build time (one warp per unit)
//   ww: own entries of water rows; pp: own entries of solute rows x the group's non-Q atoms; pw: B entries of solute rows
//   x the group's non-Q atoms.  off_*[k]: first entry of unit k (scan of the counts k_energy_counts writes).
__global__ void k_energy_counts(Dev D, const int *__restrict__ counts, int *__restrict__ n_ww, int *__restrict__ n_pp,
                                int *__restrict__ n_pw) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= D.nunit) return;
    if (u < D.ncgp_solute) {
        const int nq = D.nq_off[u + 1] - D.nq_off[u];
        n_pp[u] = nq * counts[3 * u];
        n_pw[u] = nq * counts[3 * u + 2];
    } else n_ww[u - D.ncgp_solute] = counts[3 * u];
}
__global__ void __launch_bounds__(256)
k_energy_fill(Dev D, const int *__restrict__ counts, const int *__restrict__ row_off, const uint32_t *__restrict__ rows,
              const int *__restrict__ upk, const int *__restrict__ pk_atom, const int *__restrict__ off_ww,
              const int *__restrict__ off_pp, const int *__restrict__ off_pw, int2 *__restrict__ ww_pairs,
              int2 *__restrict__ pp_pairs, int2 *__restrict__ pw_pairs) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= D.nunit) return;
    const int own = counts[3 * u], nb = counts[3 * u + 2];
    const uint32_t *r = rows + row_off[u];
    const int pi0 = upk[u];
    if (u >= D.ncgp_solute) {
        int2 *dst = ww_pairs + off_ww[u - D.ncgp_solute];
        for (int k = lane; k < own; k += 32) dst[k] = make_int2(pi0, (int)(r[k] & kIdMask));
        return;
    }
    const int nq = D.nq_off[u + 1] - D.nq_off[u];
    if (nq == 0) return;
    int2 *dpp = pp_pairs + off_pp[u];
    for (int k = lane; k < own; k += 32) {
        const uint32_t e = r[k];
        const int pb = (int)(e & kIdMask);
        const int bb = pk_atom[pb];
        const bool special = (e & kSpecialBit) != 0, same = special && D.grp_of_atom[bb] == u;
        for (int t = 0; t < nq; t++) {
            const int a = D.nq_atoms[D.nq_off[u] + t];
            uint32_t w = (uint32_t)pb | (e & (0x3fu << kImgShift));
            if (special) {
                if (a == bb) w |= 0x80000000u;
                else {
                    const int sc = special_code(D, a, bb);
                    if (sc == 0) w |= 0x80000000u;
                    else if (sc == 3) w |= 0x40000000u;
                }
                if (same && !(a < bb)) w |= 0x80000000u;   // count once inside a group (i < j, L1874)
            }
            dpp[(size_t)k * nq + t] = make_int2(pi0 + t, (int)w);
        }
    }
    int2 *dpw = pw_pairs + off_pw[u];
    const uint32_t *rb = r + own + counts[3 * u + 1];
    for (int k = lane; k < nb; k += 32) {
        const uint32_t e = rb[k];
        for (int t = 0; t < nq; t++) dpw[(size_t)k * nq + t] = make_int2(pi0 + t, (int)(e & (kIdMask | (0x3fu << kImgShift))));
    }
}

// several exclusive scans in one launch: block b scans in[b][0..n[b]) into out[b][0..n[b]] (out[n] = total)
struct ScanJobs { const int *in[8]; int *out[8]; int n[8]; };
__global__ void __launch_bounds__(1024) k_multi_scan(ScanJobs J) {
    __shared__ int warp_off[32];
    __shared__ int block_total;
    __shared__ int carry;
    const int *in = J.in[blockIdx.x];
    int *out = J.out[blockIdx.x];
    const int n = J.n[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const int v = i < n ? in[i] : 0;
        const int inc = warp_incl_scan(v, lane);
        if (lane == 31) warp_off[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int t = lane < nw ? warp_off[lane] : 0;
            const int ti = warp_incl_scan(t, lane);
            warp_off[lane] = ti - t;
            if (lane == 31) block_total = ti;
        }
        __syncthreads();
        if (i < n) out[i] = carry + warp_off[wid] + inc - v;
        __syncthreads();
        if (tid == 0) carry += block_total;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
}

}  // namespace qnb
