// pxb_tv_iter.cuh -- ONE kernel per PD3O / CondatVu iteration of a TV-type problem (K = first-order Gradient,
// 'constant' boundaries): primal half-step, dual half-step and both RelError norms in a single sweep.
//
// Why: the two-kernel form (pxb_tv_kernels.cu) moves 60 B/voxel (fp32, 3-D) because w = 2x - tau grad f - u makes a
// round trip through HBM and z is read twice.  Here a CTA owns a tile of TY rows x T2 columns and MARCHES along the
// slowest axis with a direction ("M"): per plane it computes w for its tile plus a one-sample rim (phase A), keeps
// the last four w planes in shared memory, and -- one plane behind -- finishes the dual update of its tile from
// shared memory and from the z values it still holds in registers (phase C).  HBM traffic per voxel:
//      read u, shift, z0, z1, z2  +  write u, z0, z1, z2   = 36 B      (+4 B when x is written, +4 B for x_prev)
// The update is out of place (u_in/z_in -> u_out/z_out, ping-pong buffers on the host side): the rim of a tile is
// another CTA's interior, so the old iterate must stay readable while the new one is written.
//
//   geometry      NDIR == 3: M = axis 0, rows = axis 1, columns = axis 2     (component k acts along axis k)
//                 NDIR == 2: M = axis 1, no row axis (TY == 1), columns = axis 2; axis 0 and the batch index
//                            enumerate independent images
//   work item     (image, chunk of planes along M, tile); blockIdx.x enumerates them, tiles fastest, so CTAs that
//                 run together march through the same planes and share their rims through L2
//   ring          w(m) lives in slot m & 3;  phase A of plane m+1 / m+2 never touches a slot phase C of plane m
//                 still reads, hence ONE __syncthreads per plane
//
// This file holds the algorithm, the geometry and the direct-load (L1/L2-served) form; the forms that run by default
// stage their operands by TMA: pxb_tv_tma.cuh (3-D, marching, 3-stage mbarrier ring) and pxb_tv_tile2d.cuh (2-D tiles).
//
// Like the other bodies these are __host__ __device__ so that tests/emu can run the exact per-thread code on the
// CPU (phase A for every thread, then phase C for every thread: the same order the barrier enforces).
#pragma once
#include "pxb_tv_fast.cuh"

// A/B switches for timing experiments (tools/bench_criterion.py --exp): compiled in only with -DPXB_EXPERIMENT, set per process
// through the environment variable PXB_EXP; PXB_EXP(bit) is the constant 0 in every normal build and on the host emulation.
#if defined(PXB_EXPERIMENT) && defined(__CUDA_ARCH__)
static __device__ int pxb_exp_flags;
#define PXB_EXP(bit) ((pxb_exp_flags & (bit)) != 0)
#elif defined(PXB_EXPERIMENT) && defined(__CUDACC__)
static __device__ int pxb_exp_flags;
#define PXB_EXP(bit) false
#else
#define PXB_EXP(bit) false
#endif

struct PxbIterGeom {
    int nM, nR, nC;       // extents along M, rows (1 when NDIR == 2), columns
    int64_t sM, sR;       // strides in elements (columns: 1)
    int chunk, nchunk;    // planes per work item, work items per image along M
    int ntR, ntC;         // tiles per plane
    int band;             // tile-rows per band: block order is (tile column, tile-row inside the band, chunk, band, image)
    int sub;              // images per batch item (n0 when NDIR == 2, else 1)
    int64_t sub_stride;   // elements between those images
    int64_t nimg;         // batch * sub
    int64_t vol;          // elements between components / batch items
    int ndir;
    int open_lo, open_hi; // slab cuts along M (NDIR == 3): ghost planes hold the neighbour's u, z, shift
    int64_t nblocks;
    int edge_first;       // peer-memory exchange: the first and the last chunk of every tile come early in the block order (1: first, 2: interleaved 1:3)
};

// Device-side stopping rule of iterations launched back to back without the host in the loop (pxb_pds_iter_n): the last
// thread block of an iteration evaluates RelError on the sums the iteration accumulated and, when the rule is met, raises
// ctl[0]; the launches queued behind it return at once.  ctl = {stop, done, ticket, -}: pxb_iter_ctl on the device.
struct PxbIterStop {
    int32_t* ctl;         // null: single launch, the host evaluates the criterion
    double eps_x, eps_z;  // RelError thresholds (<= 0: the variable takes no part)
    int32_t all_x, all_z; // over the batch rows: every row (1) / any row (0)   (satisfy_all, stop.py:353-382)
    int32_t table;        // bit (2*px + pz): stop for that combination of outcomes
    int32_t rows;
};

// Halo exchange fused into the iteration (pxb_pds_iter_p2p): the thread blocks that produce a slab's first / last plane store
// it into the neighbour's ghost planes as well -- peer memory over NVLink, the same 16-byte stores -- then bump a counter in the
// neighbour's memory; the neighbour's blocks that READ those ghost planes in its next iteration wait for the counter first.
// Edge chunks come first in the block order, so the data is on its way a whole iteration before it is needed and the wait is
// normally over before it starts.  No NCCL call, no side stream, one launch per iteration and rank.
template <class T>
struct PxbPeer {
    T* dn_u;                 // lower neighbour's upper ghost plane of the primal output (null: no lower neighbour)
    T* dn_z;                 // ... of z, component 0; components dn_zvol elements apart
    int64_t dn_zvol;
    T* up_z0;                // upper neighbour's lower ghost plane of z component 0
    unsigned* dn_flag;       // counters in the neighbours' memory: += 1 per thread block that has delivered its share
    unsigned* up_flag;
    const unsigned* lo_wait; // counters in this rank's memory, bumped by the lower / upper neighbour
    const unsigned* hi_wait;
    unsigned target;         // value they must have reached before this iteration reads its ghost planes
};

template <class T>
struct PxbIterPtr {
    const T* u_in;   // PD3O: u     CV: x
    const T* z_in;
    T* u_out;
    T* z_out;
    T* x_out;        // PD3O only, nullable: x is then not materialised (36 B/voxel form)
    double* norms_x; // nullable pair per batch row (RelError[x])
    double* norms_z;
    PxbIterStop stop;
    PxbPeer<T> peer;
};

#ifdef __CUDACC__
// spin (one thread) until the neighbour's counter has reached `target`; bounded, then trap: a lost signal must fail the
// launch loudly instead of hanging the device
static __device__ __forceinline__ void pxb_peer_wait(const unsigned* flag, unsigned target) {
    for (unsigned spin = 0; spin < (1u << 24); ++spin) {  // ~10 s
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - target) >= 0) return;
        __nanosleep(64);
    }
    __trap();
}
static __device__ __forceinline__ bool pxb_iter_stopped(const PxbIterStop& s) {
    return s.ctl != nullptr && *reinterpret_cast<const volatile int32_t*>(s.ctl) != 0;
}
// Called by the one thread of a block that has just added the block's sums: the block that draws the last ticket sees every
// block's contribution (fence before the ticket), evaluates the rule exactly as the host does -- sqrt(num) <= eps * sqrt(den) in
// IEEE double, the same operations NumPy performs (stop.py:371-378) -- and counts the iteration.
static __device__ __forceinline__ void pxb_iter_finish(const PxbIterStop& s, const double* nx, const double* nz, unsigned nblocks) {
    if (s.ctl == nullptr) return;
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(s.ctl + 2);
    if (atomicAdd(ticket, 1u) != nblocks - 1u) return;
    *ticket = 0u;  // (the next launch starts once this kernel has ended)
    __threadfence();
    auto met = [&](const double* n, double eps, int all) -> int {
        if (n == nullptr || !(eps > 0.0)) return 0;
        bool every = true, some = false;
        for (int r = 0; r < s.rows; ++r) {
            const double num = sqrt(__ldcg(n + 2 * r)), den = sqrt(__ldcg(n + 2 * r + 1));
            const bool ok = num <= eps * den;
            every = every && ok;
            some = some || ok;
        }
        return (all ? every : some) ? 1 : 0;
    };
    const int px = met(nx, s.eps_x, s.all_x), pz = met(nz, s.eps_z, s.all_z);
    s.ctl[1] += 1;
    if ((s.table >> (2 * px + pz)) & 1) *reinterpret_cast<volatile int32_t*>(s.ctl) = 1;
    __threadfence();
}
#endif

struct PxbIterItem {
    int64_t lin_base;  // offset of the image in u-like arrays
    int64_t z_base;    // offset of component 0 of the image in z-like arrays
    int64_t v_base;    // offset of the image inside its batch item
    int64_t b;         // batch row
    int m0, m1, r0, c0;
    int si;            // image index inside the batch item (NDIR == 2: its axis-0 coordinate)
    int full;          // the tile lies entirely inside the image (no row / column tests for its own samples)
    int nopeer;        // 1: the launch carries no peer-memory exchange at all (set by the kernel once per thread block: the per-plane tests
                       // of the three peer pointers were 4 % of the instructions); 0: test the pointers
};

template <class T, int VEC, int TXL, int TY, int NDIR>
struct PxbIterCfg {
    static constexpr bool HASR = NDIR == 3;
    static constexpr int NT = TXL * TY, T2 = TXL * VEC, RS = T2 + 2 * VEC;
    static constexpr int R0 = HASR ? 1 : 0, ROWS = HASR ? TY + 2 : TY, SLOT = ROWS * RS, NSLOT = 4;
    static constexpr int KM = 0, KR = HASR ? 1 : 0 /*unused when !HASR*/, KC = NDIR - 1;
    static constexpr size_t SMEM = sizeof(T) * NSLOT * SLOT;
};

template <class T, int VEC>
struct PxbIterThread {
    T zc[3][VEC];     // z_in of the plane phase A just visited (this thread's own samples)
    T zprev[3][VEC];  // ... of the plane before
    double acc[4];    // RelError partial sums: x (num, den), z (num, den)
};

// Block order.  CTAs that are resident together should (a) be neighbouring tiles of the SAME planes, so that the rim
// rows / columns one tile re-reads are still in L2 because its neighbour reads them at about the same time, and
// (b) move on to the next chunk of planes of the same tiles, so that the plane shared by consecutive chunks is an L2
// hit too.  Hence: a band holds as many tile-rows as there are co-resident CTAs; inside a band the order is tile
// column, tile row, then chunk.  (Measured at 1024^3: tiles-then-chunks order 5.4 TB/s, banded short chunks > 6 TB/s.)
PXB_HD PxbIterItem pxb_iter_item(const PxbIterGeom& g, int64_t blk, int ty, int t2) {
    // (the grid has fewer than 2^31 blocks -- pxb_iter_setup -- so the block arithmetic is 32-bit: the 64-bit divisions of the first
    //  version were ~1000 warp instructions per thread block)
    PxbIterItem it;
    const unsigned per_img = (unsigned)g.ntR * (unsigned)g.ntC * (unsigned)g.nchunk;
    const unsigned img = (unsigned)blk / per_img;
    unsigned rem = (unsigned)blk - img * per_img;
    int ch, tR, tC;
    // edge_first: the chunks that hold a slab's first / last plane -- what the neighbours wait for -- before the interior ones
    // edge_first == 2: ... interleaved with the interior ones, one edge block in four, over the first 4 * (edge blocks) of the grid.
    // (An A/B switch: all edge blocks at once put 296 blocks' worth of peer stores on NVLink in the first microseconds of the launch;
    // interleaved, every SM pairs an edge block with an interior one.  Measured on 2 B200s: no difference.)
    const int n_edge = g.edge_first ? (g.nchunk >= 2 ? 2 : 1) : 0;
    const unsigned tiles = (unsigned)g.ntR * (unsigned)g.ntC;
    const unsigned eb = (unsigned)n_edge * tiles;
    bool edge = rem < eb;
    unsigned idx = edge ? rem : rem - eb;
    if (g.edge_first == 2 && per_img - eb >= 3u * eb) {
        if (rem < 4u * eb) {
            edge = (rem & 3u) == 0;
            idx = edge ? (rem >> 2) : rem - (rem >> 2) - 1;
        } else {
            edge = false;
            idx = rem - eb;
        }
    }
    if (edge) {
        const unsigned e = idx / tiles, t = idx - e * tiles;
        ch = e == 0 ? 0 : g.nchunk - 1;
        tR = (int)(t / (unsigned)g.ntC);
        tC = (int)(t - (unsigned)tR * (unsigned)g.ntC);
    } else {
        rem = idx;
        const unsigned nch = (unsigned)(g.nchunk - n_edge);
        const unsigned full_band = (unsigned)g.band * (unsigned)g.ntC * nch;
        const unsigned bi = rem / full_band;
        rem -= bi * full_band;
        const int rows_here = g.ntR - (int)bi * g.band < g.band ? g.ntR - (int)bi * g.band : g.band;
        const unsigned per_chunk = (unsigned)rows_here * (unsigned)g.ntC;
        const unsigned chq = rem / per_chunk;
        const unsigned r2 = rem - chq * per_chunk;
        ch = (int)chq + (n_edge ? 1 : 0);
        const unsigned rq = r2 / (unsigned)g.ntC;
        tR = (int)bi * g.band + (int)rq;
        tC = (int)(r2 - rq * (unsigned)g.ntC);
    }
    const int64_t b = (int64_t)(img / (unsigned)g.sub);
    const int si = (int)(img - (unsigned)b * (unsigned)g.sub);
    it.b = b;
    it.si = si;
    it.v_base = (int64_t)si * g.sub_stride;
    it.lin_base = b * g.vol + it.v_base;
    it.z_base = b * g.ndir * g.vol + it.v_base;
    it.m0 = ch * g.chunk;
    it.m1 = it.m0 + g.chunk < g.nM ? it.m0 + g.chunk : g.nM;
    it.r0 = tR * ty;
    it.c0 = tC * t2;
    it.full = (it.r0 + ty <= g.nR && it.c0 + t2 <= g.nC) ? 1 : 0;
    it.nopeer = 0;
    return it;
}


// ---------------------------------------------------------------------------------------------------------
// Boundary modes other than 'constant' (numpy.pad wrap / reflect / symmetric / edge): the MODES instances of the tiled
// forms keep the 'constant' arithmetic and add (a) to K^T z of the one sample per line and side onto which the padded
// cell folds, one term from global memory (pxb_tv_fold_kz), and (b) into the cells of the w tile that lie one step
// outside the domain, w at the sample the boundary map folds them onto (pxb_tv_w_outside), because K w reads them as
// the padded array.  Nothing else changes: phase C is the 'constant' code.
// ---------------------------------------------------------------------------------------------------------
PXB_HD bool pxb_any_mode(const pxb_grad_desc& d) {
    for (int k = 0; k < d.ndir; ++k)
        if (d.mode[3 - d.ndir + k] != PXB_CONSTANT) return true;
    return false;
}

// The forms that stage their operands in shared memory (TMA) can often serve a rim cell without leaving the tile: the
// rim cell at coordinate e (= t0 - 1 or t0 + tn) of a tile [t0, t0 + tn) along an axis of length n stands for
//   e itself when it lies inside the domain,
//   the sample the boundary map folds it onto when that sample belongs to the tile (reflect / symmetric / edge on the
//   tile that touches the face; wrap only when the tile spans the axis),
// else PXB_NOSRC: a zero of the 'constant' extension, or (MODES) a far sample that pxb_tv_w_outside fetches.
PXB_HD int pxb_rim_src(int e, int n, int mode, int t0, int tn, bool modes) {
    if (e >= 0 && e < n) return e;
    if (!modes || mode == PXB_CONSTANT) return PXB_NOSRC;
    const int f = pxb_bmap(e, n, mode);
    return (f >= t0 && f < t0 + tn && f >= 0 && f < n) ? f : PXB_NOSRC;
}

// ---------------------------------------------------------------------------------------------------------
// w (and, when `store`, the new primal iterate) for W consecutive samples starting at (m, r, c).
// The samples lie inside the domain or on a ghost plane of an open slab side.
// ---------------------------------------------------------------------------------------------------------
template <class T, int W, int NDIR, int ALGO, bool NORMS, bool MODES = false>
PXB_HD void pxb_iter_w(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterItem& it, const PxbIterPtr<T>& a, int m, int r, int c,
                       bool store, T* wv, T (*zc)[W], double* acc) {
    constexpr bool HASR = NDIR == 3;
    constexpr int KM = 0, KR = 1, KC = NDIR - 1;
    const int64_t off = (int64_t)m * g.sM + (int64_t)r * g.sR + c;
    const T* __restrict__ zb = a.z_in + it.z_base + off;
    T kz[W], t[W];
    // (K_k^T z)[s] = cm*z_k[s+e] + c0*z_k[s] + cp*z_k[s-e]
    {
        const T* __restrict__ zk = zb + KM * g.vol;
        const PxbVec<T, W> cv = pxb_vload<T, W>(zk);
        if (zc) for (int j = 0; j < W; ++j) zc[KM][j] = cv.v[j];
        pxb_tv_taps_col<T, W>(zk, g.sM, cv, q.cm[KM], q.c0[KM], q.cp[KM], m > 0 || g.open_lo, m < g.nM - 1 || g.open_hi, kz);
    }
    if (HASR) {
        const T* __restrict__ zk = zb + KR * g.vol;
        const PxbVec<T, W> cv = pxb_vload<T, W>(zk);
        if (zc) for (int j = 0; j < W; ++j) zc[KR][j] = cv.v[j];
        pxb_tv_taps_col<T, W>(zk, g.sR, cv, q.cm[KR], q.c0[KR], q.cp[KR], r > 0, r < g.nR - 1, t);
        for (int j = 0; j < W; ++j) kz[j] += t[j];
    }
    {
        const T* __restrict__ zk = zb + KC * g.vol;
        const PxbVec<T, W> cv = pxb_vload<T, W>(zk);
        if (zc) for (int j = 0; j < W; ++j) zc[KC][j] = cv.v[j];
        pxb_tv_taps_row<T, W>(zk, cv, q.cm[KC], q.c0[KC], q.cp[KC], c > 0, c + W < g.nC, t);
        for (int j = 0; j < W; ++j) kz[j] += t[j];
    }
    if (MODES) pxb_tv_fold_kz<T, W, NDIR>(q, a.z_in + it.b * NDIR * g.vol, NDIR == 3 ? m : it.si, NDIR == 3 ? r : m, c, kz);
    const int64_t lin = it.lin_base + off;
    const PxbVec<T, W> old = pxb_vload<T, W>(a.u_in + lin);
    PxbVec<T, W> sh;
    for (int j = 0; j < W; ++j) sh.v[j] = T(0);
    if (q.fkind == PXB_F_SQL2) {
        if (q.shift_mode == PXB_SHIFT_LIN) sh = pxb_vload<T, W>(q.shift + lin);
        else if (q.shift_mode == PXB_SHIFT_VOL) sh = pxb_vload<T, W>(q.shift + it.v_base + off);
        else if (q.shift_mode == PXB_SHIFT_SCALAR) { for (int j = 0; j < W; ++j) sh.v[j] = q.shift[0]; }
        else if (q.shift_mode == PXB_SHIFT_MOD) { for (int j = 0; j < W; ++j) sh.v[j] = q.shift[(lin + j) % q.shift_period]; }
    }
    PxbVec<T, W> un, xo;
    if (ALGO == PXB_PD3O) {
        for (int j = 0; j < W; ++j) {
            const T x = pxb_prox_eval<T>(q.gkind, q.gp0, q.gp1, old.v[j] - q.tau * kz[j], q.tau);
            const T gf = (q.fkind == PXB_F_SQL2) ? (x + sh.v[j]) * q.two_alpha : T(0);
            const T ut = x - q.tau * gf;
            wv[j] = x + ut - old.v[j];
            un.v[j] = q.one_m_rho * old.v[j] + q.rho * ut;
            xo.v[j] = x;
        }
        if (store) {
            if (NORMS && a.norms_x) {
                const PxbVec<T, W> xprev = pxb_vload<T, W>(a.x_out + lin);
                T s0 = T(0), s1 = T(0);  // one vector's partial sums in the working precision, widened once
                for (int j = 0; j < W; ++j) {
                    const T dd = xo.v[j] - xprev.v[j];
                    s0 += dd * dd;
                    s1 += xprev.v[j] * xprev.v[j];
                }
                acc[0] += (double)s0;
                acc[1] += (double)s1;
            }
            if (a.x_out) pxb_vstore<T, W>(a.x_out + lin, xo);
        }
    } else {
        PxbVec<T, W> ga;
        if (q.fkind == PXB_F_GRADARR) ga = pxb_vload<T, W>(q.garr + lin);
        T s0 = T(0), s1 = T(0);
        for (int j = 0; j < W; ++j) {
            T gf = T(0);
            if (q.fkind == PXB_F_SQL2) gf = (old.v[j] + sh.v[j]) * q.two_alpha;
            else if (q.fkind == PXB_F_GRADARR) gf = ga.v[j];
            const T vv = old.v[j] - q.tau * gf - q.tau * kz[j];
            const T xt = pxb_prox_eval<T>(q.gkind, q.gp0, q.gp1, vv, q.tau);
            wv[j] = T(2) * xt - old.v[j];
            un.v[j] = q.rho * xt + q.one_m_rho * old.v[j];
            if (NORMS && store && a.norms_x) {
                const T dd = un.v[j] - old.v[j];
                s0 += dd * dd;
                s1 += old.v[j] * old.v[j];
            }
        }
        if (NORMS && store && a.norms_x) { acc[0] += (double)s0; acc[1] += (double)s1; }
    }
    if (store) pxb_vstore<T, W>(a.u_out + lin, un);
}

// ---------------------------------------------------------------------------------------------------------
// phase A of plane m: w(m) of the tile (+ rim on planes the work item updates) -> ring slot m & 3
// ---------------------------------------------------------------------------------------------------------
template <class T, int VEC, int TXL, int TY, int NDIR, int ALGO, bool NORMS, bool MODES = false>
PXB_HD void pxb_iter_phaseA(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterItem& it, const PxbIterPtr<T>& a, int tid, int m,
                            T* smem, PxbIterThread<T, VEC>& st) {
    using C = PxbIterCfg<T, VEC, TXL, TY, NDIR>;
    T* __restrict__ slot = smem + (m & 3) * C::SLOT;
    const bool plane_in = (m >= 0 || g.open_lo) && (m < g.nM || g.open_hi);
    const bool own = m >= it.m0 && m < it.m1;
    const int rl = tid / TXL, cx = tid - rl * TXL, cl = cx * VEC;
    // (i0, i1) of sample (m, r): NDIR == 3: (m, r);  NDIR == 2: (image index, m)
#define PXB_I0(m_, r_) (NDIR == 3 ? (m_) : it.si)
#define PXB_I1(m_, r_) (NDIR == 3 ? (r_) : (m_))
    {
        const int r = it.r0 + rl, c = it.c0 + cl;
        PxbVec<T, VEC> wv;
        if (plane_in && r < g.nR && c < g.nC) pxb_iter_w<T, VEC, NDIR, ALGO, NORMS, MODES>(q, g, it, a, m, r, c, own, wv.v, st.zc, st.acc);
        else if (MODES && (own || !plane_in)) pxb_tv_w_outside<T, VEC, NDIR, ALGO>(q, a.u_in, a.z_in, it.b, PXB_I0(m, r), PXB_I1(m, r), c, wv.v);
        else for (int j = 0; j < VEC; ++j) wv.v[j] = T(0);  // (on the planes of neighbouring chunks / slabs only the tile's own cells are read)
        pxb_vstore<T, VEC>(slot + (rl + C::R0) * C::RS + cl + VEC, wv);
    }
    if (!own) return;  // planes of the neighbouring chunks are only needed at the tile's own positions
    if (C::HASR && tid < 2 * TXL) {  // rim rows r0-1 (needed when cm != 0) and r0+TY (cp != 0):  (K w)[s] = cm w[s-e] + c0 w[s] + cp w[s+e]
        const bool top = tid < TXL;
        const T coef = top ? q.cm[C::KR] : q.cp[C::KR];
        if (coef != T(0)) {
            const int r = top ? it.r0 - 1 : it.r0 + TY, c = it.c0 + cl;
            PxbVec<T, VEC> wv;
            if (r >= 0 && r < g.nR && c < g.nC) pxb_iter_w<T, VEC, NDIR, ALGO, false, MODES>(q, g, it, a, m, r, c, false, wv.v, (T(*)[VEC]) nullptr, st.acc);
            else if (MODES) pxb_tv_w_outside<T, VEC, NDIR, ALGO>(q, a.u_in, a.z_in, it.b, PXB_I0(m, r), PXB_I1(m, r), c, wv.v);
            else for (int j = 0; j < VEC; ++j) wv.v[j] = T(0);
            pxb_vstore<T, VEC>(slot + (top ? 0 : TY + 1) * C::RS + cl + VEC, wv);
        }
    }
    if (tid >= C::NT - 2 * TY) {  // rim columns c0-1 (cm != 0) and c0+T2 (cp != 0), one sample per row
        const int h = tid - (C::NT - 2 * TY);
        const bool left = h < TY;
        const int hl = left ? h : h - TY;
        const T coef = left ? q.cm[C::KC] : q.cp[C::KC];
        if (coef != T(0)) {
            const int r = it.r0 + hl, c = left ? it.c0 - 1 : it.c0 + C::T2;
            T w1[1];
            if (r < g.nR && c >= 0 && c < g.nC) pxb_iter_w<T, 1, NDIR, ALGO, false, MODES>(q, g, it, a, m, r, c, false, w1, (T(*)[1]) nullptr, st.acc);
            else if (MODES) pxb_tv_w_outside<T, 1, NDIR, ALGO>(q, a.u_in, a.z_in, it.b, PXB_I0(m, r), PXB_I1(m, r), c, w1);
            else w1[0] = T(0);
            slot[(hl + C::R0) * C::RS + (left ? VEC - 1 : VEC + C::T2)] = w1[0];
        }
    }
#undef PXB_I0
#undef PXB_I1
}

#if defined(__CUDA_ARCH__)
// ---------------------------------------------------------------------------------------------------------
// Packed form of phase C for the headline instances (fp32, 4 samples per thread, three forward differences, L21, rho == 1):
// the same update on PAIRS of samples -- fma.rn.f32x2 / mul.rn.f32x2 (FFMA2 / FMUL2) -- with sigma folded into the tap
// coefficients.  Why: the kernel is HBM-bound in bursts but draws the board's power cap when it runs for more than a second; the SM
// clock then falls to ~1.5 GHz and its ~110 instructions per voxel no longer hide under the memory time (5.8 instead of 6.5 TB/s,
// tools/probe_sustained.py).  43 % of those instructions are fp32 arithmetic on vectors whose lanes share their coefficients.
// Results differ from the scalar body in the last bits (sigma * (c0 w + cp w') against (sigma c0) w + (sigma cp) w').
// ---------------------------------------------------------------------------------------------------------
template <int TXL, int TY, bool NORMS>
static __device__ __forceinline__ void pxb_iter_phaseC_f32x2(const PxbTvP<float>& q, const PxbIterGeom& g, const PxbIterItem& it, const PxbIterPtr<float>& a,
                                                               int tid, int mm, const float* smem, const float (*zo)[4], double* acc, float* zb_at) {
    using C = PxbIterCfg<float, 4, TXL, TY, 3>;
    const int rl = tid / TXL, cl = (tid - rl * TXL) * 4;
    const int r = it.r0 + rl, c = it.c0 + cl;
    if (!it.full && (r >= g.nR || c >= g.nC)) return;
    const int cell = (rl + C::R0) * C::RS + cl + 4;
    const float* __restrict__ s1 = smem + (mm & 3) * C::SLOT + cell;
    const float4 wc = *reinterpret_cast<const float4*>(s1);
    const float4 wm = *reinterpret_cast<const float4*>(smem + ((mm + 1) & 3) * C::SLOT + cell);  // next plane
    const float4 wr = *reinterpret_cast<const float4*>(s1 + C::RS);                               // next row
    const float hi = s1[4];                                                                      // next column of the last sample
    auto pk = [&](int i) { return *reinterpret_cast<const float2*>(q.pk[i]); };  // (c, c) pairs folded on the host
    const float2 aM2 = pk(0), aR2 = pk(1), aC2 = pk(2), bM2 = pk(3), bR2 = pk(4), bC2 = pk(5);
    const float2 wcl = make_float2(wc.x, wc.y), wch = make_float2(wc.z, wc.w);
    float2 p0l = __ffma2_rn(bM2, make_float2(wm.x, wm.y), __ffma2_rn(aM2, wcl, make_float2(zo[0][0], zo[0][1])));
    float2 p0h = __ffma2_rn(bM2, make_float2(wm.z, wm.w), __ffma2_rn(aM2, wch, make_float2(zo[0][2], zo[0][3])));
    float2 p1l = __ffma2_rn(bR2, make_float2(wr.x, wr.y), __ffma2_rn(aR2, wcl, make_float2(zo[1][0], zo[1][1])));
    float2 p1h = __ffma2_rn(bR2, make_float2(wr.z, wr.w), __ffma2_rn(aR2, wch, make_float2(zo[1][2], zo[1][3])));
    float2 p2l = __ffma2_rn(bC2, make_float2(wc.y, wc.z), __ffma2_rn(aC2, wcl, make_float2(zo[2][0], zo[2][1])));
    float2 p2h = __ffma2_rn(bC2, make_float2(wc.w, hi), __ffma2_rn(aC2, wch, make_float2(zo[2][2], zo[2][3])));
    const float2 nl = __ffma2_rn(p2l, p2l, __ffma2_rn(p1l, p1l, __fmul2_rn(p0l, p0l)));
    const float2 nh = __ffma2_rn(p2h, p2h, __ffma2_rn(p1h, p1h, __fmul2_rn(p0h, p0h)));
    const float2 sl = make_float2(fminf(1.0f, q.lam * rsqrtf(nl.x)), fminf(1.0f, q.lam * rsqrtf(nl.y)));
    const float2 sh = make_float2(fminf(1.0f, q.lam * rsqrtf(nh.x)), fminf(1.0f, q.lam * rsqrtf(nh.y)));
    p0l = __fmul2_rn(p0l, sl); p0h = __fmul2_rn(p0h, sh);
    p1l = __fmul2_rn(p1l, sl); p1h = __fmul2_rn(p1h, sh);
    p2l = __fmul2_rn(p2l, sl); p2h = __fmul2_rn(p2h, sh);
    if (NORMS) {
        float a0 = 0.f, a1 = 0.f;
        const float2 pz[3][2] = {{p0l, p0h}, {p1l, p1h}, {p2l, p2h}};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float d0 = pz[k][0].x - zo[k][0], d1 = pz[k][0].y - zo[k][1], d2 = pz[k][1].x - zo[k][2], d3 = pz[k][1].y - zo[k][3];
            a0 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            a1 += zo[k][0] * zo[k][0] + zo[k][1] * zo[k][1] + zo[k][2] * zo[k][2] + zo[k][3] * zo[k][3];
        }
        acc[2] += (double)a0;
        acc[3] += (double)a1;
    }
    float* __restrict__ zb = zb_at ? zb_at : a.z_out + it.z_base + (int64_t)mm * g.sM + (int64_t)r * g.sR + c;
    const float4 o0 = make_float4(p0l.x, p0l.y, p0h.x, p0h.y), o1 = make_float4(p1l.x, p1l.y, p1h.x, p1h.y), o2 = make_float4(p2l.x, p2l.y, p2h.x, p2h.y);
    *reinterpret_cast<float4*>(zb) = o0;
    *reinterpret_cast<float4*>(zb + g.vol) = o1;
    *reinterpret_cast<float4*>(zb + 2 * g.vol) = o2;
    if (it.nopeer) return;
    if (mm == 0 && a.peer.dn_z != nullptr) {  // peer-memory exchange: the first owned plane of every component goes down ...
        float* pd = a.peer.dn_z + (int64_t)r * g.sR + c;
        *reinterpret_cast<float4*>(pd) = o0;
        *reinterpret_cast<float4*>(pd + a.peer.dn_zvol) = o1;
        *reinterpret_cast<float4*>(pd + 2 * a.peer.dn_zvol) = o2;
    }
    if (mm == g.nM - 1 && a.peer.up_z0 != nullptr) *reinterpret_cast<float4*>(a.peer.up_z0 + (int64_t)r * g.sR + c) = o0;  // ... the last one of component 0 up
}
#endif

// ---------------------------------------------------------------------------------------------------------
// phase C of plane mm: z_out(mm) = (1-rho) z + rho prox_{sigma h*}(z + sigma K w) for the tile, w from the ring,
// z (old) from `zo` (registers).
// ---------------------------------------------------------------------------------------------------------
template <class T, int VEC, int TXL, int TY, int NDIR, bool NORMS, class S = PxbSpecAny>
PXB_HD void pxb_iter_phaseC(const PxbTvP<T>& q, const PxbIterGeom& g, const PxbIterItem& it, const PxbIterPtr<T>& a, int tid, int mm,
                            const T* smem, const T (*zo)[VEC], double* acc, T* zb_at = nullptr) {
    // zb_at: z_out at this thread's samples on plane mm, component 0, when the caller keeps a running pointer (saves the 64-bit
    // index arithmetic per plane: 19 of the 463 warp instructions per plane of the TMA form)
    using C = PxbIterCfg<T, VEC, TXL, TY, NDIR>;
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4 && VEC == 4 && NDIR == 3 && S::SCHEME == PXB_SCHEME_FWD && S::HK == PXB_DUAL_L21) {
        if (q.rho1 && !PXB_EXP(64)) {  // (uniform)
            pxb_iter_phaseC_f32x2<TXL, TY, NORMS>(q, g, it, a, tid, mm, smem, zo, acc, zb_at);
            return;
        }
    }
#endif
    const int rl = tid / TXL, cx = tid - rl * TXL, cl = cx * VEC;
    const int r = it.r0 + rl, c = it.c0 + cl;
    if (!it.full && (r >= g.nR || c >= g.nC)) return;
    const int cell = (rl + C::R0) * C::RS + cl + VEC;
    const T* __restrict__ s1 = smem + (mm & 3) * C::SLOT + cell;
    const PxbVec<T, VEC> wc = pxb_vload<T, VEC>(s1);
    T p[NDIR][VEC];
    {  // along M: neighbouring ring slots
        T kw[VEC];
        for (int j = 0; j < VEC; ++j) kw[j] = q.c0[C::KM] * wc.v[j];
        if (pxb_has_cm<S>(q, C::KM)) {
            const PxbVec<T, VEC> n = pxb_vload<T, VEC>(smem + ((mm - 1) & 3) * C::SLOT + cell);
            for (int j = 0; j < VEC; ++j) kw[j] += q.cm[C::KM] * n.v[j];
        }
        if (pxb_has_cp<S>(q, C::KM)) {
            const PxbVec<T, VEC> n = pxb_vload<T, VEC>(smem + ((mm + 1) & 3) * C::SLOT + cell);
            for (int j = 0; j < VEC; ++j) kw[j] += q.cp[C::KM] * n.v[j];
        }
        for (int j = 0; j < VEC; ++j) p[C::KM][j] = zo[C::KM][j] + q.sigma * kw[j];
    }
    if (C::HASR) {  // along the rows of the slot
        T kw[VEC];
        for (int j = 0; j < VEC; ++j) kw[j] = q.c0[C::KR] * wc.v[j];
        if (pxb_has_cm<S>(q, C::KR)) {
            const PxbVec<T, VEC> n = pxb_vload<T, VEC>(s1 - C::RS);
            for (int j = 0; j < VEC; ++j) kw[j] += q.cm[C::KR] * n.v[j];
        }
        if (pxb_has_cp<S>(q, C::KR)) {
            const PxbVec<T, VEC> n = pxb_vload<T, VEC>(s1 + C::RS);
            for (int j = 0; j < VEC; ++j) kw[j] += q.cp[C::KR] * n.v[j];
        }
        for (int j = 0; j < VEC; ++j) p[C::KR][j] = zo[C::KR][j] + q.sigma * kw[j];
    }
    {  // along the row (rim cells hold w of the neighbouring tile, or 0 outside the domain)
        T lo = T(0), hi = T(0);
        if (pxb_has_cm<S>(q, C::KC)) lo = s1[-1];
        if (pxb_has_cp<S>(q, C::KC)) hi = s1[VEC];
        for (int j = 0; j < VEC; ++j) {
            T kw = q.c0[C::KC] * wc.v[j];
            if (pxb_has_cp<S>(q, C::KC)) kw += q.cp[C::KC] * (j + 1 < VEC ? wc.v[j + 1 < VEC ? j + 1 : 0] : hi);
            if (pxb_has_cm<S>(q, C::KC)) kw += q.cm[C::KC] * (j > 0 ? wc.v[j > 0 ? j - 1 : 0] : lo);
            p[C::KC][j] = zo[C::KC][j] + q.sigma * kw;
        }
    }
    T a0 = T(0), a1 = T(0);  // RelError[z] partial sums of this vector in the working precision, widened once below
    for (int j = 0; j < VEC; ++j) {
        T grp[PXB_MAX_DIRS];
        for (int k = 0; k < NDIR; ++k) grp[k] = p[k][j];
        if (!PXB_EXP(16)) pxb_dual_prox_group<T>(pxb_hkind<S>(q), NDIR, q.lam, q.sigma, grp);
        for (int k = 0; k < NDIR; ++k) p[k][j] = grp[k];
    }
    if (!q.rho1)  // (uniform; rho == 1: (1 - rho) z + rho p is p, bit for bit)
        for (int k = 0; k < NDIR; ++k)
            for (int j = 0; j < VEC; ++j) p[k][j] = q.one_m_rho * zo[k][j] + q.rho * p[k][j];
    if (NORMS && !PXB_EXP(4))
        for (int k = 0; k < NDIR; ++k)
            for (int j = 0; j < VEC; ++j) {
                const T dd = p[k][j] - zo[k][j];
                a0 += dd * dd;
                a1 += zo[k][j] * zo[k][j];
            }
    T* __restrict__ zb = zb_at ? zb_at : a.z_out + it.z_base + (int64_t)mm * g.sM + (int64_t)r * g.sR + c;
    // peer-memory exchange: the first owned plane of every component goes down, the last one of component 0 goes up
    // (the plane tests first: they are uniform and almost never true)
    const bool dn = NDIR == 3 && !it.nopeer && mm == 0 && a.peer.dn_z != nullptr, up = NDIR == 3 && !it.nopeer && mm == g.nM - 1 && a.peer.up_z0 != nullptr;
    for (int k = 0; k < NDIR; ++k) {
        PxbVec<T, VEC> o;
        for (int j = 0; j < VEC; ++j) o.v[j] = p[k][j];
        pxb_vstore<T, VEC>(zb + k * g.vol, o);
        if (dn) pxb_vstore<T, VEC>(a.peer.dn_z + k * a.peer.dn_zvol + (int64_t)r * g.sR + c, o);
        if (up && k == 0) pxb_vstore<T, VEC>(a.peer.up_z0 + (int64_t)r * g.sR + c, o);
    }
    if (NORMS) { acc[2] += (double)a0; acc[3] += (double)a1; }
}

// The marching loop of one thread between barriers (`sync` is __syncthreads on the device; the host emulation
// runs the two phases as separate sweeps over the threads instead of calling this).
//   lag = 1 when (K w)[m] needs w[m+1] (cp != 0 along M): phase C then trails phase A by one plane.
struct PxbIterRange {
    int mlo, mhi, lag;
};
template <class T>
PXB_HD PxbIterRange pxb_iter_range(const PxbTvP<T>& q, const PxbIterItem& it) {
    PxbIterRange R;
    R.lag = q.cp[0] != T(0) ? 1 : 0;
    R.mlo = it.m0 - (q.cm[0] != T(0) ? 1 : 0);
    R.mhi = it.m1 + R.lag;
    return R;
}

// ---------------------------------------------------------------------------------------------------------
// host side: eligibility + geometry (shared by the launcher and by tests/emu)
// ---------------------------------------------------------------------------------------------------------
// returns 0 when the single-kernel iteration applies, else a reason code (> 0)
// `allow_modes`: folding boundary modes are served by the MODES instances.  On a slab an open side is not a boundary; a
// closed one folds onto the slab's own planes, which 'wrap' along axis 0 cannot do unless the slab is the whole volume
// (the caller opens every side for a ring exchange instead).
inline int pxb_iter_setup(const pxb_grad_desc& d, const pxb_pds_params& P, int vec, int ty, int t2, int chunk_hint, int resident, PxbTvCoef& cf,
                          PxbIterGeom& g, bool allow_modes = false) {
    if (!pxb_tv_fast_coefs(d, cf)) return 1;
    if (d.ndir != 2 && d.ndir != 3) return 2;
    if (P.hkind != PXB_DUAL_L21 && P.hkind != PXB_DUAL_L1) return 3;
    if (pxb_any_mode(d) && !allow_modes) return 4;
    if (d.ndir == 3 && d.mode[0] == PXB_WRAP && d.slab.open_lo != d.slab.open_hi) return 4;
    if (d.shape[2] % vec) return 5;
    if (d.shape[0] < 1 || d.shape[1] < 1 || d.shape[2] < 1 || d.batch < 1) return 9;
    const PxbGeom gg = pxb_geom(d.shape);
    g.ndir = d.ndir;
    g.vol = pxb_vol(gg, d.slab);
    g.nC = gg.n2;
    if (d.ndir == 3) {
        g.nM = gg.n0; g.nR = gg.n1; g.sM = gg.s0; g.sR = gg.s1;
        g.sub = 1; g.sub_stride = 0;
        g.open_lo = d.slab.open_lo; g.open_hi = d.slab.open_hi;
        // a ghost w plane needs z one plane further out when the stencil along M is two-sided
        const bool two_sided = cf.cm[0] != 0.0 && cf.cp[0] != 0.0;
        if ((g.open_lo || g.open_hi) && d.slab.halo < (two_sided ? 2 : 1)) return 6;
    } else {
        if (d.slab.halo != 0 || d.slab.open_lo || d.slab.open_hi) return 7;
        g.nM = gg.n1; g.nR = 1; g.sM = gg.s1; g.sR = 0;
        g.sub = gg.n0; g.sub_stride = gg.s0;
        g.open_lo = g.open_hi = 0;
    }
    g.nimg = d.batch * g.sub;
    g.ntR = (g.nR + ty - 1) / ty;
    g.ntC = (g.nC + t2 - 1) / t2;
    const int64_t tiles = (int64_t)g.ntR * g.ntC * g.nimg;
    g.band = resident / g.ntC;  // `resident` = CTAs of this kernel the GPU holds at once (148 SMs x CTAs per SM)
    if (g.band < 1) g.band = 1;
    if (g.band > g.ntR) g.band = g.ntR;
    int chunk = chunk_hint;
    if (chunk <= 0) {  // short chunks keep co-resident CTAs on the same planes (their rims stay L2 hits); with 8 planes
        chunk = g.nM < 8 ? g.nM : 8;  // 1/8 of the planes is read twice, through L2.  Measured 1024^3: 8 -> 6.2 TB/s, 128 -> 5.3 TB/s
    }
    if (chunk > g.nM) chunk = g.nM;
    g.chunk = chunk;
    g.nchunk = (g.nM + chunk - 1) / chunk;
    g.nblocks = tiles * g.nchunk;
    g.edge_first = 0;
    if (g.nblocks <= 0 || g.nblocks > 0x7fffffffLL) return 8;
    return 0;
}
