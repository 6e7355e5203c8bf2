// pxb_stencil3d_fast.cu -- the single-pass separable 3-D stencil (design: pxb_stencil3d.cuh) for the PSF-shaped case: K x K x K taps,
// K in {3, 5, 7, 9}, centred along the rows of the array.  Same marching scheme and register ring as k_stencil3d, rebuilt around
// what its profile showed (profiles/r02_f_ncu_full_k_stencil3d.csv: 68 instructions per voxel, 65 % issue-active at 41 % of the
// copy bandwidth, i.e. ISSUE-bound -- only 42 % of those instructions were FMAs):
//   * every extent is a template parameter: the three passes are fully unrolled, the coefficients are constant-bank operands of
//     the FMAs (a T-typed copy in the parameter block), the alignment padding of the row factor (its leading zero taps) is skipped
//     at compile time instead of multiplied through;
//   * the in-plane passes are swapped: columns first (box -> `mid`, on the TX + 2*VEC columns of the box: 6 % halo work), rows
//     second (`mid` -> registers, own samples only) -- the other order filters the K-1 halo ROWS too (38 % extra at TY = 16);
//   * fp32: the column pass and the K-tap combination along axis 0 act on all samples of a vector with ONE coefficient: they run
//     on packed pairs (fma.rn.f32x2 -> FFMA2), half the instructions;
//   * `mid` is double-buffered (one barrier per plane instead of two), three box stages (two planes of prefetch);
//   * addresses advance by pointer increments (the 64-bit index arithmetic of the epilogue operand was 15 % of the instructions).
// Everything else -- other extents, periodic epilogue operands -- stays with k_stencil3d.
#include "pxb_launch.cuh"
#include "pxb_tma_util.cuh"
#include "pxb_stencil3d.cuh"

#include <type_traits>

namespace {

template <class T, int K>
struct St3FastP {
    T c0[K], c1[K], c2[K];      // factors along axes 0, 1, 2 (c2: the K source taps, without the alignment padding)
    unsigned long long c0d[K], c1d[K];  // fp32: (c, c) pairs -- FFMA2 takes them as uniform-register operands straight from the constant bank
    T alpha, beta;
    unsigned long long alphad, betad;
    const T* add;               // dense (batch, n0, n1, n2) or null
    int n0, n1, n2;
    int64_t batch, vol;
    int c0i, c1i;               // centres along axes 0, 1
    int c2p;                    // centre along axis 2 after padding (a multiple of VEC): the box starts at x0 - c2p
    int lo_planes, hi_planes, chunk, nchunk, ntx, nty;
};

template <class T, int VEC, int K>
struct St3FastCfg {
    static constexpr int TXL = 32, TX = TXL * VEC, TY = 16, R = 2, NT = TXL * (TY / R);
    static constexpr int C2 = K / 2, EXTRA = (VEC - C2 % VEC) % VEC;          // leading zero taps of the padded row factor
    static constexpr int NV = (K + EXTRA - 1 + 2 * VEC - 1) / VEC;           // vectors a thread's row window spans
    static constexpr int BW = TX + (NV - 1) * VEC, BH = TY + K - 1;
    static constexpr int BOX = (BH * BW + 31) / 32 * 32, MID = TY * BW;
    static constexpr int NBOX = 3;
    static constexpr int CP = (C2 + EXTRA) / VEC;                            // halo vector columns left of the tile (the box starts at x0 - CP*VEC)
    static constexpr int HCOLS = (NV - 1) * VEC, HALO = HCOLS * TY;          // halo columns of the box; column-pass outputs there, ONE SAMPLE per thread
    static_assert(HALO <= NT, "one halo sample per thread at most");
    static constexpr size_t SMEM = sizeof(T) * (NBOX * BOX + 2 * MID);
};

// packed pairs (fp32): acc += c * v on two adjacent samples with one instruction
struct P2 {
    unsigned long long u;
};
static __device__ __forceinline__ P2 p2_make(float a, float b) {
    P2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.u) : "f"(a), "f"(b));
    return r;
}
static __device__ __forceinline__ void p2_split(P2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.u)); }
static __device__ __forceinline__ P2 p2_fma(unsigned long long a, P2 b, P2 c) {
    P2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a), "l"(b.u), "l"(c.u));
    return r;
}
static __device__ __forceinline__ P2 p2_mul(unsigned long long a, P2 b) {
    P2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a), "l"(b.u));
    return r;
}

// a vector of VEC samples as the unit the column pass / the axis-0 combination work on; `cd` = the coefficient as a (c, c) pair
template <class T, int VEC>
struct Acc;
template <>
struct Acc<float, 4> {
    P2 a, b;
    __device__ __forceinline__ void zero() { a = p2_make(0.f, 0.f); b = a; }
    __device__ __forceinline__ void load(const float* p) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        a = p2_make(t.x, t.y); b = p2_make(t.z, t.w);
    }
    __device__ __forceinline__ void fma(float, unsigned long long cd, const Acc& v) { a = p2_fma(cd, v.a, a); b = p2_fma(cd, v.b, b); }
    __device__ __forceinline__ void scale(float, unsigned long long cd) { a = p2_mul(cd, a); b = p2_mul(cd, b); }
    __device__ __forceinline__ void store(float* p) const {
        float4 t;
        p2_split(a, t.x, t.y); p2_split(b, t.z, t.w);
        *reinterpret_cast<float4*>(p) = t;
    }
    __device__ __forceinline__ void from(const float* v) { a = p2_make(v[0], v[1]); b = p2_make(v[2], v[3]); }
};
template <>
struct Acc<double, 2> {
    double x, y;
    __device__ __forceinline__ void zero() { x = y = 0.0; }
    __device__ __forceinline__ void load(const double* p) {
        const double2 t = *reinterpret_cast<const double2*>(p);
        x = t.x; y = t.y;
    }
    __device__ __forceinline__ void fma(double c, unsigned long long, const Acc& v) { x += c * v.x; y += c * v.y; }
    __device__ __forceinline__ void scale(double c, unsigned long long) { x *= c; y *= c; }
    __device__ __forceinline__ void store(double* p) const { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
    __device__ __forceinline__ void from(const double* v) { x = v[0]; y = v[1]; }
};

template <class T, int VEC, int K, bool ADD>
__global__ void __launch_bounds__(256, 2) k_stencil3d_fast(const __grid_constant__ St3FastP<T, K> p, const __grid_constant__ CUtensorMap map, T* __restrict__ out) {
    using C = St3FastCfg<T, VEC, K>;
    using V = Acc<T, VEC>;
    extern __shared__ __align__(128) unsigned char pxb_st3f_smem[];
    __shared__ __align__(8) uint64_t bar[C::NBOX];
    T* boxes = reinterpret_cast<T*>(pxb_st3f_smem);
    T* mids = boxes + C::NBOX * C::BOX;
    const int tid = threadIdx.x;
    unsigned blk = blockIdx.x;
    const int tx = blk % (unsigned)p.ntx; blk /= (unsigned)p.ntx;
    const int ty = blk % (unsigned)p.nty; blk /= (unsigned)p.nty;
    const int ch = blk % (unsigned)p.nchunk;
    const int b = blk / (unsigned)p.nchunk;
    const int x0 = tx * C::TX, y0 = ty * C::TY;
    const int m0 = ch * p.chunk, m1 = min(p.n0, m0 + p.chunk);
    const int pl_lo = m0 - p.c0i, pl_hi = m1 + K - 1 - p.c0i;                      // input planes this chunk needs
    const int ra = max(pl_lo, -p.lo_planes), rb = min(pl_hi, p.n0 + p.hi_planes);  // ... those that exist
    constexpr uint32_t bytes = (uint32_t)(C::BH * C::BW * sizeof(T));
    auto issue = [&](int pl) {
        const int s = (pl - ra) % C::NBOX;
        mbar_expect_tx(&bar[s], bytes);
        tma_load_4d(boxes + s * C::BOX, &map, &bar[s], x0 - p.c2p, y0 - p.c1i, pl + p.lo_planes, b);
    };
    if (tid == 0) {
        for (int s = 0; s < C::NBOX; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int pl = ra; pl < ra + C::NBOX && pl < rb; ++pl) issue(pl);

    V ring[K][C::R];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int r = 0; r < C::R; ++r) ring[k][r].zero();
    const int xl = (tid & 31) * VEC, yl = (tid >> 5) * C::R;
    const int x = x0 + xl, y = y0 + yl;
    const bool col_ok = x < p.n2;
    const bool full_tile = x0 + C::TX <= p.n2 && y0 + C::TY <= p.n1;
    const int64_t plane = (int64_t)p.n1 * p.n2;
    // output / epilogue-operand pointers of this thread's first row on plane q = pl_lo - (K - 1 - c0) ... advanced by one plane per step
    const int q0 = pl_lo - (K - 1 - p.c0i);
    T* optr = out + (int64_t)b * p.vol + (int64_t)q0 * plane + (int64_t)y * p.n2 + x;
    const T* aptr = ADD ? p.add + ((int64_t)b * p.n0 + q0) * plane + (int64_t)y * p.n2 + x : nullptr;
    // column pass, work per thread: the tile's own vector column `lane` of rows yl, yl+1 (8 warps x 32 lanes = the 256 items of
    // the tile) plus ONE SAMPLE of the halo columns (threads < HALO).  Dealing the halo out as vector items gave one warp a second
    // round while seven waited at the barrier (25 % of the stall samples); a ninth warp for them costs the second CTA its registers.
    const int own_off = yl * C::BW + C::CP * VEC + xl;
    const int hrow = tid / C::HCOLS, hcol0 = tid - hrow * C::HCOLS;
    const int halo_off = hrow * C::BW + (hcol0 < C::CP * VEC ? hcol0 : hcol0 + C::TX);
    const bool halo_item = tid < C::HALO;
    int s = 0;            // box stage of the next plane that exists
    uint32_t par = 0;
    int mslot = 0;

    // one plane: `u` = ring slot (compile-time after unrolling); CLEAN = the plane exists, its output plane is emitted, the tile is full
    auto step = [&](const int u, const int pl, auto clean_tag) {
        constexpr bool CLEAN = decltype(clean_tag)::value;
        const bool have = CLEAN || (pl >= ra && pl < rb);
        const int q = pl - (K - 1 - p.c0i);
        const bool emit = CLEAN || q >= m0;
        PxbVec<T, VEC> addv[C::R];  // (kept as loaded: packing them here would wait for the loads before the passes)
        if (ADD && emit && (CLEAN || col_ok)) {  // in flight during the passes
#pragma unroll
            for (int r = 0; r < C::R; ++r)
                if (CLEAN || y + r < p.n1) addv[r] = pxb_vload<T, VEC>(aptr + r * p.n2);
        }
        if (have) {
            const T* box = boxes + s * C::BOX;
            T* mid = mids + mslot * C::MID;
            mbar_wait(&bar[s], par);
            {   // column pass: box -> mid
                const T* src = box + own_off;
                V acc[C::R];
#pragma unroll
                for (int r = 0; r < C::R; ++r) acc[r].zero();
#pragma unroll
                for (int i = 0; i < C::R + K - 1; ++i) {
                    V v;
                    v.load(src + i * C::BW);
#pragma unroll
                    for (int r = 0; r < C::R; ++r)
                        if (i - r >= 0 && i - r < K) acc[r].fma(p.c1[i - r], p.c1d[i - r], v);
                }
#pragma unroll
                for (int r = 0; r < C::R; ++r) acc[r].store(mid + own_off + r * C::BW);
                if (halo_item) {
                    const T* hs = box + halo_off;
                    T a = T(0);
#pragma unroll
                    for (int t = 0; t < K; ++t) a += p.c1[t] * hs[t * C::BW];
                    mid[halo_off] = a;
                }
            }
            __syncthreads();  // `mid` complete; this stage's box is free
            if (tid == 0 && pl + C::NBOX < rb) issue(pl + C::NBOX);
            // row pass: own samples, window of NV vectors per row; the padding taps (EXTRA leading zeros) are skipped
#pragma unroll
            for (int r = 0; r < C::R; ++r) {
                T w[C::NV * VEC];
                const T* src = mid + (yl + r) * C::BW + xl;
#pragma unroll
                for (int n = 0; n < C::NV; ++n) {
                    const PxbVec<T, VEC> t = pxb_vload<T, VEC>(src + n * VEC);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) w[n * VEC + j] = t.v[j];
                }
                T o[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) o[j] = T(0);
#pragma unroll
                for (int t = 0; t < K; ++t)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) o[j] += p.c2[t] * w[C::EXTRA + t + j];
                ring[u][r].from(o);
            }
            s = s + 1 == C::NBOX ? 0 : s + 1;
            if (s == 0) par ^= 1u;
            mslot ^= 1;
        } else {
#pragma unroll
            for (int r = 0; r < C::R; ++r) ring[u][r].zero();
        }
        if (emit && (CLEAN || col_ok)) {
#pragma unroll
            for (int r = 0; r < C::R; ++r) {
                if (CLEAN || y + r < p.n1) {
                    V a;
                    a.zero();
#pragma unroll
                    for (int k = 0; k < K; ++k) a.fma(p.c0[k], p.c0d[k], ring[(u + 1 + k) % K][r]);
                    a.scale(p.alpha, p.alphad);
                    if (ADD) {
                        V av;
                        av.from(addv[r].v);
                        a.fma(p.beta, p.betad, av);
                    }
                    a.store(optr + r * p.n2);
                }
            }
        }
        optr += plane;
        if (ADD) aptr += plane;
    };

    for (int base = pl_lo; base < pl_hi; base += K) {
        // a group of K planes that all exist, all emit and lie in a full tile runs without per-plane tests (every group of a
        // chunk but the first and the last)
        const bool clean = full_tile && base >= ra && base + K <= rb && base + K <= pl_hi && base - (K - 1 - p.c0i) >= m0;
        if (clean) {
#pragma unroll
            for (int u = 0; u < K; ++u) step(u, base + u, std::true_type{});
        } else {
#pragma unroll
            for (int u = 0; u < K; ++u)
                if (base + u < pl_hi) step(u, base + u, std::false_type{});
        }
    }
}

template <class T, int VEC, int K>
int launch(const PxbSt3P& g, const void* in, T* out, cudaStream_t s, cudaError_t* err) {
    using C = St3FastCfg<T, VEC, K>;
    if (g.s.bw != C::BW || g.s.bh != C::BH || g.s.extra != C::EXTRA) return 31;  // (geometry the generic set-up derived differently)
    St3FastP<T, K> p;
    auto dup = [](T v) {  // (c, c) as one 64-bit value (fp32 only; unused in fp64)
        unsigned long long r = 0;
        if (sizeof(T) == 4) {
            uint32_t w;
            memcpy(&w, &v, 4);
            r = ((unsigned long long)w << 32) | w;
        }
        return r;
    };
    for (int k = 0; k < K; ++k) {
        p.c0[k] = T(g.coef0[k]);
        p.c1[k] = T(g.s.coef1[k]);
        p.c2[k] = T(g.s.coef2[k + C::EXTRA]);
        p.c0d[k] = dup(p.c0[k]);
        p.c1d[k] = dup(p.c1[k]);
    }
    p.alpha = T(g.s.alpha); p.beta = T(g.s.beta);
    p.alphad = dup(p.alpha); p.betad = dup(p.beta);
    p.add = (const T*)g.s.add;
    p.n0 = g.n0; p.n1 = g.s.n1; p.n2 = g.s.n2;
    p.batch = g.batch; p.vol = g.vol;
    p.c0i = g.c0; p.c1i = g.s.c1; p.c2p = g.s.c2;
    p.lo_planes = g.lo_planes; p.hi_planes = g.hi_planes; p.chunk = g.chunk; p.nchunk = g.nchunk; p.ntx = g.s.ntx; p.nty = g.s.nty;
    const int64_t s0 = (int64_t)g.s.n1 * g.s.n2;
    const uint64_t dim[4] = {(uint64_t)g.s.n2, (uint64_t)g.s.n1, (uint64_t)(g.n0 + g.lo_planes + g.hi_planes), (uint64_t)g.batch};
    const uint64_t stride[4] = {1, (uint64_t)g.s.n2, (uint64_t)s0, (uint64_t)g.vol};
    const uint32_t box[4] = {(uint32_t)C::BW, (uint32_t)C::BH, 1, 1};
    alignas(64) CUtensorMap map;
    if (!pxb_tma_encode_cached<T>(4, (const T*)in - (int64_t)g.lo_planes * s0, dim, stride, box, &map)) return 32;
    const unsigned grid = (unsigned)((int64_t)g.s.ntx * g.s.nty * g.nchunk * g.batch);
    auto go = [&](auto k) {
        cudaError_t e = pxb_smem_attr_once((const void*)k, (int)C::SMEM);
        if (e != cudaSuccess) { *err = e; return; }
        k<<<grid, C::NT, C::SMEM, s>>>(p, map, out);
        *err = cudaGetLastError();
    };
    if (p.add) go(k_stencil3d_fast<T, VEC, K, true>);
    else go(k_stencil3d_fast<T, VEC, K, false>);
    return 0;
}

template <class T>
int pick(const PxbSt3P& g, const void* in, void* out, cudaStream_t s, cudaError_t* err) {
    constexpr int VEC = 16 / (int)sizeof(T);
    switch (g.k0) {
        case 3: return launch<T, VEC, 3>(g, in, (T*)out, s, err);
        case 5: return launch<T, VEC, 5>(g, in, (T*)out, s, err);
        case 7: return launch<T, VEC, 7>(g, in, (T*)out, s, err);
        case 9: return launch<T, VEC, 9>(g, in, (T*)out, s, err);
        default: return 30;
    }
}

}  // namespace

// 0: launched (or *err set); > 0: not this kernel's case.  `g` is the parameter block after pxb_st3_setup.
int pxb_st3_fast_try(int dtype, const PxbSt3P& g, const void* in, void* out, cudaStream_t s, cudaError_t* err) {
    const int K = g.k0;
    if (g.s.k1 != K || g.s.k2src != K || g.s.c2 - g.s.extra != K / 2) return 30;  // K x K x K, centred along the rows
    if (g.s.add != nullptr && g.s.add_period > 0) return 33;                      // periodic epilogue operand
    return dtype == PXB_F32 ? pick<float>(g, in, out, s, err) : pick<double>(g, in, out, s, err);
}
