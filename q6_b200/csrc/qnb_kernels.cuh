// qnb_kernels.cuh -- sm_100a kernels of the Qdyn6 nonbonded path.
//
// Design (DESIGN.md has the long form):
//  * "units" = solute charge groups [0,ncgp_solute) followed by water molecules.  Pair lists
//    are held at unit level as FULL (symmetric) neighbour rows: every unit owns a row with
//    all partners it has a listed pair with, so each force kernel accumulates forces only
//    on the atoms of its own unit -- no scatter atomics, deterministic.  The reference's
//    half list (checkerboard rule, nonbondene.f90:1855) survives as the "owner" bit of an
//    entry: the side that the reference would hold the pair on computes its energy, and
//    qnb_export_list() expands exactly those entries.
//  * cut-off tests are FP64 on switch atoms, literally the reference's expression, so lists
//    are exact.  Pair forces are FP32 on coordinates taken relative to the row's own origin
//    (FP64 subtraction first), LJ energies FP32 with FP64 accumulation, Coulomb ENERGIES in
//    FP64 (rsqrt seed + one Halley step) because neutral-group sums cancel to ~1e-7 of their
//    terms and the parity bar is 1e-6 on the net value.
//  * Q-atom kernels (per-state, softcore, lambda-weighted) and LRF are FP64 throughout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda/std/tuple>

namespace qnb {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxStatesDev = 8;
constexpr int kRowWarps = 4;    // warps (of one block) that share a unit's row in the force kernels
constexpr int kESlots = 32;     // energy accumulators are replicated to keep same-address atomics rare
// Row entry: bits 0-23 packed atom index, 24-29 periodic image of the pair (any-atom mode), 30 special, 31 owner
constexpr uint32_t kOwnerBit = 0x80000000u;    // this row's unit is the reference's "i" side of the pair
constexpr uint32_t kSpecialBit = 0x40000000u;  // solute partner atom with an excluded/1-4/self relation
constexpr uint32_t kIdMask = 0x00ffffffu;
constexpr int kImgShift = 24;
constexpr int kMaxPacked = 1 << 24;

struct QPar4 { double A, B, el, score; };

struct Grid {
    double org[3];
    double inv_cell[3];
    double inv_box[3];   // periodic: 1/L
    int n[3];
    int ncell;
    int periodic;
};

// Static device tables + per-build state.  Passed by value to kernels.
struct Dev {
    int natom, nat_solute, nwat, ncgp, ncgp_solute, nunit, nqat, nstates, nct;
    int use_PBC, use_LRF, geometric, spc_water, qswitch0;
    int any_atom;   // iuse_switch_atom == 0: any-atom charge-group cut-offs (nb??lis2*)
    int sharded;    // the i-ranges do not cover everything: ownership of a pair must be checked
    int shard_rows; // sharded builds partition ROWS (and LRF targets), not pairs: see row_in_shard
    double el14;
    float el14f;
    double box[3], inv_box[3];
    double xpcent[3];
    // shard (1-based inclusive, as calculation_assignment)
    int pp_s, pp_e, pw_s, pw_e, qp_s, qp_e, ww_s, ww_e, qw_s, qw_e, at_s, at_e;
    // atoms
    const double *crg;          // [natom]
    const float *crgf;          // [natom]
    const int *ctype;           // [natom]
    const uint8_t *is_q, *excl, *qbonded;
    const int *grp_of_atom;
    // groups
    const int *g_first, *g_n, *g_switch, *g_atoms, *g_nq;   // g_nq: non-Q atoms per group
    // units
    const int *u_sw, *u_grp;
    const uint8_t *u_excl;
    // LJ tables
    const float *ljf;           // [nct][3][2] (a,b) by code
    const double *ljd;          // same in FP64 (energies)
    const uint8_t *ljcode;      // [nct][nct]
    // specials
    const int *sp_off, *sp_partner;
    const uint8_t *sp_code;
    const int *gs_off, *gs_atoms;
    const int *nq_off, *nq_atoms;   // non-Q atoms per solute group
    // water sites
    float wq[3];                // site charges
    double wqd[3];
    int wct[3];
    float wwA[9], wwB[9], wwQ[9];
    double wwQd[9], wwAd[9], wwBd[9];
    // solute non-Q atom compaction (rows of the solute kernel are per group)
    // Q
    const int *iqseq;           // [nqat]
    const QPar4 *qp_tab;        // [(iq*nstates+s)][nat_solute]
    const QPar4 *qw_tab;        // [(iq*nstates+s)][3]
    const float4 *qp_tabf, *qw_tabf;   // same as (A, B, el, score) in FP32: partner-side gradients
};

struct Cut {
    double rc2[3];      // pp, pw, ww
    double rclrf2;
    int lrf_all[3];     // "no LRF cut-off" sentinel per class (box builders)
    double rcq2;
    double Rq;
    double rmax2;   // any-atom mode: twice the largest switch-atom-to-atom distance of a solute group
    // kernel parameters live in constant memory: select by class without dynamic indexing (which would force a
    // local-memory copy of the struct)
    __host__ __device__ double rc2_of(int cls) const { return cls == 0 ? rc2[0] : cls == 1 ? rc2[1] : rc2[2]; }
    __host__ __device__ int lrf_all_of(int cls) const { return cls == 0 ? lrf_all[0] : cls == 1 ? lrf_all[1] : lrf_all[2]; }
};

// ------------------------------------------------------------------ one launch for several systems
// The step kernels are written as Body functors (struct XBody { static __device__ void run(args..., BX, NBX, BY, NBY) })
// with a thin __global__ wrapper each.  k_batched runs the same body for W systems (lambda windows of a FEP farm) in ONE
// launch: blockIdx.z names the system, its argument tuple lies in device memory and is copied to shared memory first,
// its own grid size comes from `grids` (blocks beyond it leave at once).  A 12 k-atom system cannot fill 148 SMs with a
// single wave of any kernel; seven of them per launch amortise the launch, the ramp and the tail.
template <class Body, int THREADS, int MINB, class... P>
__global__ void __launch_bounds__(THREADS, MINB) k_batched(const cuda::std::tuple<P...> *__restrict__ args, const int2 *__restrict__ grids) {
    using T = cuda::std::tuple<P...>;
    __shared__ alignas(16) unsigned char raw[sizeof(T)];
    const int2 g = grids[blockIdx.z];
    if ((int)blockIdx.x >= g.x || (int)blockIdx.y >= g.y) return;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(args + blockIdx.z);
        uint32_t *dst = reinterpret_cast<uint32_t *>(raw);
        for (int i = threadIdx.x; i < (int)(sizeof(T) / 4); i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const T &t = *reinterpret_cast<const T *>(raw);
    cuda::std::apply([&](const P &...p) { Body::run(p..., (int)blockIdx.x, g.x, (int)blockIdx.y, g.y); }, t);
}

// ------------------------------------------------------------------ several kernels in one launch
// A CUDA graph pays ~2.5 us of front-end time per kernel node: the step of a 12 k-atom system (pack + seven kernels of
// 8-14 us each on their own streams) took 32 us whatever the grids (r03f) and 25 us even for 600 atoms (r02y).  k_uber
// runs several Body functors in ONE launch: the grid is the concatenation of the bodies' grids, a block looks up which
// body owns its index.  Arguments travel by value (constant bank), exactly as in the single kernels.
template <class F> struct BodyRunArgs;
template <class... P> struct BodyRunArgs<void (*)(P...)> { using type = cuda::std::tuple<cuda::std::decay_t<P>...>; };
template <class Tup, size_t... I>
cuda::std::tuple<cuda::std::tuple_element_t<I, Tup>...> body_args_head(cuda::std::index_sequence<I...>);
// the argument tuple of Body::run without its trailing (BX, NBX, BY, NBY)
template <class Body>
using body_args_t = decltype(body_args_head<typename BodyRunArgs<decltype(&Body::run)>::type>(
    cuda::std::make_index_sequence<cuda::std::tuple_size<typename BodyRunArgs<decltype(&Body::run)>::type>::value - 4>{}));
template <class Body> struct USlot { body_args_t<Body> a; int gx, gy; };   // gx * gy blocks; gx == 0: body absent this step
template <class Body>
__device__ __forceinline__ bool uber_try(const USlot<Body> &s, int &b) {
    const int n = s.gx * s.gy;
    if (b < n) {
        const int bx = b % s.gx, by = b / s.gx;
        cuda::std::apply([&](const auto &...p) { Body::run(p..., bx, s.gx, by, s.gy); }, s.a);
        return true;
    }
    b -= n;
    return false;
}
template <int MINB, class... Body>
__global__ void __launch_bounds__(128, MINB) k_uber(const __grid_constant__ USlot<Body>... s) {
    int b = blockIdx.x;
    (void)(uber_try(s, b) || ...);   // the first body whose block range holds b runs
}

// ------------------------------------------------------------------ all-reduce over peer memory
// The ranks of one node (one process per GPU) map each other's arena [out | lrf | control] through CUDA IPC; NVSwitch
// gives every GPU full bandwidth to every peer, so the sum is ONE kernel per rank: the owner of slice r (1/n of the
// elements) reads that slice from every rank, adds the n values in rank order and writes the sum back into EVERY
// rank's buffer -- 2 (n-1)/n of the buffer crosses NVLink per rank, nothing is staged, every rank ends up with
// bit-identical sums.  Two flag barriers in peer memory bracket it (all partial results complete / all sums delivered and
// nobody still reading); the epoch lives in device memory, so the kernel can sit in a CUDA graph.
constexpr int kMaxPeers = 8;
struct P2PComm {
    double *base[kMaxPeers];   // arena of every rank as mapped in THIS process (base[rank] = own)
    size_t ctl_off;            // doubles from the arena base to the control words
    int rank, n;
};
// control words (unsigned) of a rank: [0] epoch, [1] blocks finished, [2] error, [16 + p] "rank p has arrived", [32 + p]
// "rank p is done"
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// wait until *p >= target (wrap-safe); gives up after ~4 s of GPU clock and flags the error instead of hanging the device
__device__ __forceinline__ bool spin_until(const unsigned *p, unsigned target, unsigned *err) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - target) < 0) {
        // a barrier that failed once fails fast from then on (qnb_comm_status reports it)
        if (*(volatile unsigned *)err || clock64() - t0 > 8000000000ll) { atomicExch(err, 1u); return false; }
        __nanosleep(64);
    }
    return true;
}
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// phase stamps of the last call (ns, %globaltimer), control words 64..71 as u64[4]: entry, all ranks arrived, own slice
// delivered (block 0), all ranks done -- read by qnb_bench_allreduce_phases
__global__ void __launch_bounds__(256)
k_p2p_allreduce(P2PComm C, size_t off, size_t count, unsigned *ctl) {
    unsigned long long *stamp = reinterpret_cast<unsigned long long *>(ctl + 64);
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp[0] = gtime_ns();
    const unsigned epoch = ld_acquire_sys(ctl) + 1u;
    // ---- everybody's partial results are complete (this kernel follows the rank's step kernels in stream order)
    if (blockIdx.x == 0 && threadIdx.x < C.n) {
        unsigned *peer = reinterpret_cast<unsigned *>(C.base[threadIdx.x] + C.ctl_off);
        __threadfence_system();
        st_release_sys(peer + 16 + C.rank, epoch);
    }
    if (threadIdx.x < C.n) spin_until(ctl + 16 + threadIdx.x, epoch, ctl + 2);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp[1] = gtime_ns();
    // ---- own slice: sum over the ranks in rank order, deliver to every rank (16-byte accesses)
    const size_t n2 = (count + 1) / 2;                       // double2 elements (the arena parts are padded)
    const size_t per = (n2 + C.n - 1) / C.n;
    const size_t lo = per * C.rank, hi = lo + per < n2 ? lo + per : n2;
    for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
#pragma unroll
        for (int p = 0; p < kMaxPeers; p++)
            if (p < C.n) {
                const double2 v = reinterpret_cast<const double2 *>(C.base[p] + off)[i];
                s.x += v.x; s.y += v.y;
            }
#pragma unroll
        for (int p = 0; p < kMaxPeers; p++)
            if (p < C.n) reinterpret_cast<double2 *>(C.base[p] + off)[i] = s;
    }
    // ---- the last block of this rank tells everybody, waits for everybody, and opens the next epoch.  One system-scope
    // fence per block, by its leader after the block barrier (the pattern of a grid-wide sync): a fence in every thread
    // put ~18 k membar.sys on the memory system at once
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp[2] = gtime_ns();
    __shared__ unsigned last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(ctl + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence_system();   // the other blocks' deliveries (fenced before their count) precede the "done" below
    if (threadIdx.x < C.n) {
        unsigned *peer = reinterpret_cast<unsigned *>(C.base[threadIdx.x] + C.ctl_off);
        st_release_sys(peer + 32 + C.rank, epoch);
        spin_until(ctl + 32 + threadIdx.x, epoch, ctl + 2);
    }
    __syncthreads();
    if (threadIdx.x == 0) { stamp[3] = gtime_ns(); ctl[1] = 0u; __threadfence(); st_release_sys(ctl, epoch); }
}

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
// Sum N per-lane values over the warp with N + O(log N) shuffles instead of 5 N: at every butterfly level the two
// partner lanes split the remaining values, each keeps one half and hands the other over.  After the levels with
// masks 16..2 a lane holds ONE value; split_slot() tells which (or -1).  FP32: the inputs are FP32 partial sums.
template <int N, int MASK>
__device__ __forceinline__ float split_reduce(const float (&v)[N], int lane) {
    if constexpr (N == 1) {
        float s = v[0];
#pragma unroll
        for (int o = MASK; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        return s;
    } else {
        static_assert(MASK >= 2, "split_reduce: at most 16 values");
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & MASK) != 0;
        float w[H];
#pragma unroll
        for (int i = 0; i < H; i++) {
            const float hi = (H + i < N) ? v[H + i] : 0.f;
            const float send = up ? v[i] : hi, keep = up ? hi : v[i];
            w[i] = keep + __shfl_xor_sync(kFull, send, MASK);
        }
        return split_reduce<H, MASK / 2>(w, lane);
    }
}
// index of the value split_reduce<N,16> leaves in this lane, -1 for lanes that hold a duplicate or a padding slot
template <int N, int MASK>
__device__ __forceinline__ int split_slot(int lane, int base = 0, int count = N) {
    if constexpr (N == 1) return (count >= 1 && (lane & (2 * MASK - 1)) == 0) ? base : -1;
    else {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & MASK) != 0;
        return split_slot<H, MASK / 2>(lane, up ? base + H : base, up ? count - H : min(count, H));
    }
}
// double -> float ONCE: nvcc rematerialises a plain (float) cast of a loop-invariant double inside the loop to save a
// register (r02o: 4.7 F2F per interaction in k_lrf_allpairs, the 16-lane XU pipe 31 % busy with them); a volatile asm
// cannot be duplicated, so the converted value stays in its register
__device__ __forceinline__ float f32_once(double v) {
    float r;
    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(v));
    return r;
}
// MUFU.RSQ alone (rsqrtf() wraps it in denormal scaling: three more instructions per pair)
__device__ __forceinline__ float rsqrt_fast(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// boxlength*q_nint(shift*inv_boxl): round() = half away from zero = Fortran nint
__device__ __forceinline__ double pshift(double v, double L, double invL) {
    return __dmul_rn(L, round(__dmul_rn(v, invL)));
}
// qvec_square without contraction: (x*x + y*y) + z*z
__device__ __forceinline__ double sq3(double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

// 1/sqrt(r2) in FP64 from an FP32 seed: one Halley step, relative error ~ seed_err^3
__device__ __forceinline__ double rsqrt_refine(double r2, float seed) {
    double y = (double)seed;
    double e = fma(-r2 * y, y, 1.0);
    double t = fma(0.375, e, 0.5) * e;
    return fma(y, t, y);
}

__device__ __forceinline__ double rinv_f64(double r2) { return rsqrt_refine(r2, rsqrt_fast((float)r2)); }

__device__ __forceinline__ int special_code(const Dev &D, int a, int b) {
    // -1: ordinary pair; kPairExcluded(0): skip; 3: 1-4 pair
    int lo = D.sp_off[a], hi = D.sp_off[a + 1];
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (D.sp_partner[m] < b) lo = m + 1; else hi = m;
    }
    if (lo < D.sp_off[a + 1] && D.sp_partner[lo] == b) return (int)D.sp_code[lo];
    return -1;
}


// class of a unit pair and the reference's owner side.  Units: [0,ns) solute groups, then waters.
// returns class 0 pp, 1 pw, 2 ww; owner_is_u: the reference lists the pair while looping i = u.
__device__ __forceinline__ int pair_class(int u, int v, int ns, bool &owner_is_u) {
    const bool us = u < ns, vs = v < ns;
    if (us != vs) { owner_is_u = us; return 1; }            // pw: the solute group is always "i" (L2864)
    const int a = us ? u + 1 : u - ns + 1, b = us ? v + 1 : v - ns + 1;   // 1-based group / water numbers
    if (a == b) owner_is_u = true;
    else {
        // kept while looping ig=a iff not ((a>b and even sum) or (a<b and odd sum)) (L1855-1857)
        const bool even = ((a + b) & 1) == 0;
        owner_is_u = (a > b) ? !even : even;
    }
    return us ? 0 : 2;
}
__device__ __forceinline__ bool in_shard(const Dev &D, int cls, int owner_unit) {
    if (cls == 2) { int w = owner_unit - D.ncgp_solute + 1; return w >= D.ww_s && w <= D.ww_e; }
    int g = owner_unit + 1;
    return cls == 0 ? (g >= D.pp_s && g <= D.pp_e) : (g >= D.pw_s && g <= D.pw_e);
}
// Row partition of a sharded build.  The reference gives rank r the pairs whose outer-loop unit i lies in r's range and
// lets r add both sides of the pair; gather_nonbond then sums d over the ranks.  Since the sum is all that survives,
// any partition of the (unit, partner) work gives the same d, E and LRF moments.  Here a rank builds the COMPLETE row
// (own, mirror and other-kind entries) of every unit of its ranges and nothing else: the candidate scan of the list
// build is sharded too (with the pair partition every rank screened all unit pairs), the gradient of an atom is
// complete on exactly one rank, every pair's energy is counted once (on the rank that holds its owner entry), and the
// owner entries a rank holds are still exactly the reference's list for calculation_assignment's i-range.
//   solute row u: pp partners if u in the pp range, water partners (pw, owner side) if u in the pw range
//   water row u:  water partners (ww) and solute partners (water side of pw) if u in the ww range
__device__ __forceinline__ bool row_in_shard(const Dev &D, int cls, int u) {
    if (u >= D.ncgp_solute) { const int w = u - D.ncgp_solute + 1; return w >= D.ww_s && w <= D.ww_e; }
    const int g = u + 1;
    return cls == 0 ? (g >= D.pp_s && g <= D.pp_e) : (g >= D.pw_s && g <= D.pw_e);
}
__device__ __forceinline__ bool row_in_any_shard(const Dev &D, int u) {
    return row_in_shard(D, 0, u) || row_in_shard(D, 1, u);
}
// squared switch-atom distance exactly as the builders compute it
__device__ __forceinline__ double unit_r2(const Dev &D, const double *pu, const double *pv) {
    // Explicitly rounded operations: no FMA contraction, so r2 is bit-identical to the reference's
    // a%x**2 + a%y**2 + a%z**2 (math.f90:206) and borderline pairs fall on the same side.
    double dx, dy, dz;
    if (!D.use_PBC) {
        // q_dist4(x(is),x(ja)) = |x(ja)-x(is)|^2 (math.f90:254)
        dx = __dsub_rn(pv[0], pu[0]); dy = __dsub_rn(pv[1], pu[1]); dz = __dsub_rn(pv[2], pu[2]);
    } else {
        // shift = x(is)-x(ja); r2 = q_dist4(shift, boxlength*q_nint(shift*inv_boxl)) (L1991-1992)
        double sx = __dsub_rn(pu[0], pv[0]), sy = __dsub_rn(pu[1], pv[1]), sz = __dsub_rn(pu[2], pv[2]);
        dx = __dsub_rn(pshift(sx, D.box[0], D.inv_box[0]), sx);
        dy = __dsub_rn(pshift(sy, D.box[1], D.inv_box[1]), sy);
        dz = __dsub_rn(pshift(sz, D.box[2], D.inv_box[2]), sz);
    }
    return sq3(dx, dy, dz);
}

// periodic image triple {-1,0,1}^3 <-> 6 bits (0 -> 0, 1 -> +1, 2 -> -1 per component)
__device__ __forceinline__ uint32_t img_pack(int nx, int ny, int nz) {
    auto c = [](int n) { return (uint32_t)(n == 0 ? 0 : n > 0 ? 1 : 2); };
    return c(nx) | (c(ny) << 2) | (c(nz) << 4);
}
__device__ __forceinline__ uint32_t img_negate(uint32_t code) {
    // swap 1 <-> 2 in every 2-bit field
    const uint32_t nz = (code | (code >> 1)) & 0x15u;   // fields that are non-zero
    return code ^ (nz * 3u);
}
__device__ __forceinline__ double img_comp(uint32_t entry, int d) {
    const uint32_t c = (entry >> (kImgShift + 2 * d)) & 3u;
    return c == 0 ? 0.0 : c == 1 ? 1.0 : -1.0;
}

// FP32 screening of a switch-atom distance against a squared cut-off: -1 surely inside, +1 surely outside, 0 too
// close to call (the FP64 test decides).  Positions are |x| < ~200 A in FP32: |error(r2)| < 1e-3*r2 + 0.05.
__device__ __forceinline__ int screen_r2(float r2f, float cut2) {
    if (r2f < cut2 * (1.0f - 1e-3f) - 0.05f) return -1;
    if (r2f > cut2 * (1.0f + 1e-3f) + 0.05f) return 1;
    return 0;
}
__device__ __forceinline__ float screen_dist2(const Dev &D, float dx, float dy, float dz) {
    if (D.use_PBC) {
        dx -= (float)D.box[0] * rintf(dx * (float)D.inv_box[0]);
        dy -= (float)D.box[1] * rintf(dy * (float)D.inv_box[1]);
        dz -= (float)D.box[2] * rintf(dz * (float)D.inv_box[2]);
    }
    return dx * dx + dy * dy + dz * dz;
}

// Outcome of the reference's cut-off test for the unit pair (i-unit = the reference's outer-loop group).
struct PairTest { bool listed, lrf; uint32_t img; };

// i-unit io, j-unit jo in OWNER orientation (io is what the reference loops as ig / iw).  pi, pj: switch-atom positions.
__device__ __forceinline__ PairTest unit_pair_test(const Dev &D, const Cut &C, const double *__restrict__ x, int cls,
                                                   int io, int jo, const double *pi, const double *pj) {
    PairTest t;
    t.img = 0;
    if (!D.any_atom || cls == 2) {
        // switching atoms (nbpplist*, nbpwlist*, nbwwlist*): one distance decides both branches
        const double r2 = unit_r2(D, pi, pj);
        t.listed = r2 <= C.rc2_of(cls);
        t.lrf = !t.listed && (r2 <= C.rclrf2 || C.lrf_all_of(cls));
        return t;
    }
    // any atom (nbpplis2*, nbpwlis2*): group io = solute group; jo = solute group (pp) or water (pw, first atom only)
    const double rc2 = C.rc2_of(cls);
    const int fi = D.g_first[io], ni = D.g_n[io];
    int fj, nj;
    if (cls == 0) { fj = D.g_first[jo]; nj = D.g_n[jo]; } else { fj = 0; nj = 1; }
    const int jwat = D.nat_solute + 3 * (jo - D.ncgp_solute);
    bool inside = false, inside_lrf = false;
    double r2 = 1.0;
    for (int a = 0; a < ni && !inside; a++) {
        const int i = D.g_atoms[fi + a];
        for (int b = 0; b < nj && !inside; b++) {
            const int j = cls == 0 ? D.g_atoms[fj + b] : jwat;
            if (!D.use_PBC) {
                r2 = sq3(__dsub_rn(x[3 * j], x[3 * i]), __dsub_rn(x[3 * j + 1], x[3 * i + 1]), __dsub_rn(x[3 * j + 2], x[3 * i + 2]));
                if (r2 <= rc2) inside = true;
            } else {
                const double sx = __dsub_rn(x[3 * i], x[3 * j]), sy = __dsub_rn(x[3 * i + 1], x[3 * j + 1]),
                             sz = __dsub_rn(x[3 * i + 2], x[3 * j + 2]);
                const double nx = round(__dmul_rn(sx, D.inv_box[0])), ny = round(__dmul_rn(sy, D.inv_box[1])),
                             nz = round(__dmul_rn(sz, D.inv_box[2]));
                r2 = sq3(__dsub_rn(__dmul_rn(D.box[0], nx), sx), __dsub_rn(__dmul_rn(D.box[1], ny), sy),
                         __dsub_rn(__dmul_rn(D.box[2], nz), sz));
                if (r2 <= rc2) { inside = true; t.img = img_pack((int)nx, (int)ny, (int)nz); }
                else if (r2 <= C.rclrf2 || C.rclrf2 == -1.0) inside_lrf = true;   // squared argument vs -one (L1283, L2437)
            }
        }
    }
    t.listed = inside;
    // sphere: r2 of the LAST pair examined decides the LRF branch (L1499, L2601); box: any examined pair inside RcLRF
    t.lrf = !inside && (D.use_PBC ? inside_lrf : (r2 <= C.rclrf2));
    return t;
}

}  // namespace qnb
