// Fast path of the symmetry-function evaluation for sm_100a (double precision, values + central gradients).
//
// Same mathematics as acsf.cu (reference pantea/descriptors/acsf/acsf.py:163-330, angular.py:51-66, cutoff.py:74-79),
// specialised for the RuNNer-style potentials the benchmark uses -- one tanhu cutoff class per element, every angular
// group a single G3 member with integer zeta, every neighbour type in at most one group of its centre, no minimum image
// on r_jk (box >= 4 rc) -- and rebuilt around the instruction count of the triplet loop, which is what bounds the
// kernel (FP64 issue):
//
//  * per-neighbour weights.  exp(-eta (r_j^2 + r_k^2 + r_jk^2)) fc_j fc_k fc_jk = W_j W_k fc_jk exp(-eta r_jk^2) with
//    W_n = sqrt(pref) fc(r_n) exp(-eta r_n^2) staged once per neighbour, and the radial part of the gradient collapses
//    into Q_n = fc'(r_n)/fc(r_n) - 2 eta r_n;
//  * r_jk^2 = r_j^2 + r_k^2 - 2 r_j r_k cos(theta) from the unit-vector dot product the angular factor needs anyway
//    (records hold sqrt(2) r so that the product needs no extra factor);
//  * one cubic-convergence step after the MUFU seeds of the reciprocal root and the reciprocal;
//  * a 256-entry 2^(i/256) table so that the two exponentials need a degree-4 polynomial;
//  * 80-byte neighbour records [u_x u_y | u_z r^2 | 1/r sqrt2 r | W Q | fc q] read with four LDS.128 per role
//    (20-bank stride: conflict-free for consecutive records), pair-list entries that hold the two record byte offsets;
//  * pair lists padded per group to a multiple of 32 entries with an entry that points at the all-zero record behind
//    the last neighbour, staged 16 bytes per lane with cp.async: the loop has no bounds, validity or tail handling.
//
// The pair filter of this path tests 32 x 32 tiles of neighbour pairs on the tensor cores (TF32 mma.sync, inclusive
// margin; see mma_tf32_16x8x8 below), keeps a 32-bit survivor mask per lane and emits a tile's entries with one warp
// scan, in an order that spreads the evaluation's record reads over the shared-memory banks.
#include "acsf_common.cuh"

namespace pantea {

constexpr int kRec2 = 10;                 // doubles per staged neighbour record
constexpr int kRec2Bytes = kRec2 * 8;     // 80
constexpr int kTab2 = 256;                // entries of the 2^(i/256) table
#ifndef PANTEA_CHUNK2
#define PANTEA_CHUNK2 8
#endif
constexpr int kChunk2 = PANTEA_CHUNK2;    // pair-list iterations per cp.async group (4 or 8)
#ifndef PANTEA_FILTER2_WARPS
#define PANTEA_FILTER2_WARPS 4  // with 7 blocks per SM (72 registers): 28 warps; 8 x 3 at 80 registers measured 6 % slower
#endif
constexpr int kFilter2Warps = PANTEA_FILTER2_WARPS;

// Atoms per block of the evaluation / filter kernels: the warps of a block draw the block's atoms from a shared counter
// one after the other.  With one atom per warp the block lives as long as its most expensive atom (an O centre walks
// twice the triplets of an H centre) while the other warps' slots idle: 32 % of the warp time of the water benchmark.
#ifndef PANTEA_ATOMS_PER_WARP
#define PANTEA_ATOMS_PER_WARP 4
#endif
constexpr int kAtomsPerWarp = PANTEA_ATOMS_PER_WARP;
#ifndef PANTEA_FILTER_ATOMS_PER_WARP
#define PANTEA_FILTER_ATOMS_PER_WARP 1  // the filter's atoms differ less in cost; dynamic assignment measured slower
#endif
constexpr int kFilterAtomsPerWarp = PANTEA_FILTER_ATOMS_PER_WARP;

__device__ __forceinline__ int next_item(int* counter, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return __shfl_sync(kFullMask, v, 0);
}


#ifndef PANTEA_NU2
#define PANTEA_NU2 2
#endif
constexpr int kPadTo2f = 32 * PANTEA_NU2;  // list segments are padded to whole loop iterations of the evaluation

__host__ __device__ inline size_t eval2_smem_bytes(int scap, int n_sf) {
    return ((size_t)(scap + 1) * kRec2Bytes + (size_t)n_sf * 4 * 8 + 15) & ~size_t(15);
}

// ---- scalar helpers -------------------------------------------------------------------------------------------------
// Constants of the triplet loop live in constant memory so that they fold into the DFMA operands (c[bank][offset]);
// written as literals ptxas re-materialises them with two MOVs per use once the register budget is reached.
static __constant__ double kK2[8] = {
    0x1.71547652b82fep+8,    // [0] 256 / ln 2
    6755399441055744.0,      // [1] 1.5 * 2^52
    -0x1.62e42fee00000p-9,   // [2] -ln2/256 with the low 21 mantissa bits zero: n * kK2[2] is exact
    -0x1.a39ef35793c76p-41,  // [3] -(ln2/256 - hi)
    4.16666666666666666667e-02, 1.66666666666666666667e-01,  // [4], [5] 1/24, 1/6
    0.0, 0.0};

// sqrt(a) for a > 0: MUFU reciprocal-root seed (>= 20 good bits) and one step of cubic convergence,
// a y (1 + h/2 + 3 h^2/8) with h = 1 - a y^2, arranged so that every constant is an FP64 immediate
__device__ __forceinline__ double sqrt_cubic(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double ay = a * y;
    const double h = fma(-ay, y, 1.0);
    const double u = h * 0.5, v = h * 0.75;
    return fma(ay, fma(u, v, u), ay);
}
// 1/a: MUFU seed, one cubic step r += r (e + e^2), e = 1 - a r
__device__ __forceinline__ double rcp_cubic(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    const double e = fma(-a, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

// exp(y) = 2^k T[i] p(f), n = round(y 256/ln2) = 256 k + i, f = y - n ln2/256, |f| <= ln2/512, p of degree 4
// (truncation f^5/120 < 4e-17).  |y| < 700 guaranteed by the caller.
__device__ __forceinline__ double exp_tab256(double y, const double* __restrict__ tab) {
    const double t = fma(y, kK2[0], kK2[1]);
    const int n = __double2loint(t);
    const double nf = t - kK2[1];
    double f = fma(nf, kK2[2], y);
    f = fma(nf, kK2[3], f);
    double p = fma(f, kK2[4], kK2[5]);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
#ifdef PANTEA_EXP_PROBE
    p *= tab[n & 15];  // timing probe only (wrong values): conflict-free table reads
#else
    p *= tab[n & (kTab2 - 1)];
#endif
    return __hiloint2double(__double2hiint(p) + (n >> 8) * 0x100000, __double2loint(p));
}
// max(x, ~0) for finite x through the sign / exponent word: negative values (and -0) come out as a positive number
// below 1e-308 + (hi >= `floor_hi`); one integer instruction instead of the NaN-aware FP64 maximum
__device__ __forceinline__ double clamp_hi(double x, int floor_hi) {
    return __hiloint2double(max(__double2hiint(x), floor_hi), __double2loint(x));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

struct __align__(16) D2 { double a, b; };
__device__ __forceinline__ D2 lds128(const unsigned char* p) { return *reinterpret_cast<const D2*>(p); }

// ------------------------------------------------------------------------------------------------
// pair filter
// ------------------------------------------------------------------------------------------------
#ifndef PANTEA_EXACT_ON
#define PANTEA_EXACT_ON 1
#endif
constexpr int kStrip2 = 1024 + 128;  // entries of a warp's emission strip: one full tile behind an unflushed remainder

// The pair test runs on the tensor cores.  For a resident neighbour l and a swept neighbour s (vectors from the centre)
//     (|s - l|^2 - thr) / 2  =  -l.s + |s|^2 / 2 + (|l|^2 / 2 - thr / 2)
// is one row-times-column product of a 16 x 8 x 8 TF32 tile: row l of A = (-x, -y, -z, 1, L1, L2, 0, 0) with the row
// constants L1 = |l|^2 / 2 - thr / 2 (cutoff test) and L2 = |l|^2 - r2max / 2 (Gaussian screening); column s of B =
// (x, y, z, |s|^2 / 2, 1, 0, 0, 0) for the cutoff test and (x, y, z, |s|^2, 0, 1, 0, 0) for the screening test.  A pair is
// kept when the result is negative (sign bit).  The staged vectors are rounded to TF32 first, so the products are exact in
// the FP32 accumulator and the test is the exact distance of the ROUNDED points (each moved by at most 2^-11 of its
// length) plus the TF32 rounding / truncation of the three constants: the threshold is inflated by 5e-3 of the squared list
// radius to cover both.  The lists are inclusive anyway (the evaluation's cutoff function makes the final cut); the margin adds
// ~0.7 % zero-weight entries.  One instruction replaces 128 x 9 scalar ones.
__device__ __forceinline__ float to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}
// Survivor masks of a 32 x 32 tile follow the accumulator layout: lane (g = lane / 4, t = lane % 4) holds, for column
// block c (8 swept neighbours) and row block h (16 residents), elements j = 2 u + p at tile row 16 h + 8 u + g and tile
// column 2 t + p of the block.  Bit 31 - (8 c + 4 h + j) of the lane's mask is that element, i.e. byte 3 - c belongs to
// column block c.  Which neighbour sits in a tile row / column is free: row (h, u, g) is resident 4 g + 2 h + u and column
// n of block c is swept neighbour 8 c + (n + 2 c) % 8, so that the entries a lane emits combine 4 consecutive residents
// with 8 swept neighbours of distinct index modulo 8 -- the evaluation's 16-byte record reads of consecutive entries then
// fall into different shared-memory banks (records 8 apart share their banks).
__host__ __device__ constexpr int tile_row(int g, int h, int u) { return 4 * g + 2 * h + u; }
__host__ __device__ constexpr int tile_col(int t, int c, int p) { return 8 * c + ((2 * t + p + 2 * c) & 7); }
struct TriMaskTable {
    unsigned m[32];
    constexpr TriMaskTable() : m() {
        for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            unsigned v = 0;
            for (int c = 0; c < 4; ++c)
                for (int h = 0; h < 2; ++h)
                    for (int j = 0; j < 4; ++j)
                        if (tile_col(t, c, j & 1) < tile_row(g, h, j >> 1)) v |= 1u << (31 - (8 * c + 4 * h + j));
            m[lane] = v;
        }
    }
};
__constant__ const TriMaskTable kTriMask2 = TriMaskTable();
// rows below n (indexed [n][g]) / columns below n (indexed [n][t]) of a tile, n = 0 .. 32
struct RowMaskTable {
    unsigned m[33][8];
    constexpr RowMaskTable() : m() {
        for (int n = 0; n <= 32; ++n)
            for (int g = 0; g < 8; ++g) {
                unsigned v = 0;
                for (int c = 0; c < 4; ++c)
                    for (int h = 0; h < 2; ++h)
                        for (int j = 0; j < 4; ++j)
                            if (tile_row(g, h, j >> 1) < n) v |= 1u << (31 - (8 * c + 4 * h + j));
                m[n][g] = v;
            }
    }
};
struct ColMaskTable {
    unsigned m[33][4];
    constexpr ColMaskTable() : m() {
        for (int n = 0; n <= 32; ++n)
            for (int t = 0; t < 4; ++t) {
                unsigned v = 0;
                for (int c = 0; c < 4; ++c)
                    for (int h = 0; h < 2; ++h)
                        for (int j = 0; j < 4; ++j)
                            if (tile_col(t, c, j & 1) < n) v |= 1u << (31 - (8 * c + 4 * h + j));
                m[n][t] = v;
            }
    }
};
__constant__ const RowMaskTable kRowMask2 = RowMaskTable();
__constant__ const ColMaskTable kColMask2 = ColMaskTable();  // swept index below the resident one (lower triangle of a chunk's own tile)

// State of one warp's walk over the angular groups of its atom.
struct Filter2 {
    const float4* sf4;   // staged TF32-rounded (dx, dy, dz, r^2 / 2) of the neighbours
    int32_t* list;       // the atom's pair list in global memory
    int32_t* strip;      // the warp's shared-memory emission strip
    int pair_cap, off, fill, lane;
    int rb;      // bytes per staged neighbour record of the evaluation that will walk the list (80: double, 48: single)
    float thr1;  // inclusive squared cutoff of r_jk (margins: see above)
    float r2max;  // Gaussian screening: pairs with r_j^2 + r_k^2 + r_jk^2 above it are dropped (unused without screening)
    bool cls_test, screen, exact;
    int slot;  // exact coincidence test (only when the binning saw atoms at the same position)

    // the 8 bits of one column block (MSB first: rows 2 h + u = 0 .. 3, two columns each) -> entries at q
    __device__ __forceinline__ void emit_block(int32_t*& q, unsigned m8, const int (&rpart)[4], int e0, int e1) const {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (m8 & (0x80u >> k)) { *q = rpart[k >> 1] + ((k & 1) ? e1 : e0); ++q; }
    }

    // compacts `mask` (tile layout above; rpart[2 h + u]: entry part of the lane's four rows, cb0 + tile_col * stride:
    // entry part of its columns) into the strip, lane by lane, then flushes the strip's full 128-entry blocks with one
    // 16-byte store per lane
    __device__ __forceinline__ void emit(unsigned mask, int ncb, const int (&rpart)[4], int cb0, int stride) {
        const int t2 = 2 * (lane & 3);
        const int c = __popc(mask);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(kFullMask, incl, 31);
        int32_t* q = strip + fill + (incl - c);
        if (ncb == 4) {  // full tile: the four column blocks are independent store chains
            int32_t* q1 = q + __popc(mask >> 24);
            int32_t* q2 = q + __popc(mask >> 16);
            int32_t* q3 = q + __popc(mask >> 8);
            const int e0 = cb0 + t2 * stride, e1 = cb0 + (8 + ((t2 + 2) & 7)) * stride;
            const int e2 = cb0 + (16 + ((t2 + 4) & 7)) * stride, e3 = cb0 + (24 + ((t2 + 6) & 7)) * stride;
            emit_block(q, mask >> 24, rpart, e0, e0 + stride);
            emit_block(q1, mask >> 16, rpart, e1, e1 + stride);
            emit_block(q2, mask >> 8, rpart, e2, e2 + stride);
            emit_block(q3, mask, rpart, e3, e3 + stride);
        } else {
            for (int cb = 0; cb < ncb; ++cb) {
                const int e0 = cb0 + (8 * cb + ((t2 + 2 * cb) & 7)) * stride;
                emit_block(q, mask >> (24 - 8 * cb), rpart, e0, e0 + stride);
            }
        }
        fill += total;
        __syncwarp();
        const int nblk = fill >> 7;
        if (nblk > 0) {
            if (off + nblk * 128 <= pair_cap)  // warp-uniform; an overflowing list is only counted
                for (int blk = 0; blk < nblk; ++blk)
                    *reinterpret_cast<int4*>(list + off + blk * 128 + 4 * lane) = *reinterpret_cast<const int4*>(strip + blk * 128 + 4 * lane);
            const int rem = fill & 127;
            int4 keep = make_int4(0, 0, 0, 0);
            if (4 * lane < rem) keep = *reinterpret_cast<const int4*>(strip + nblk * 128 + 4 * lane);
            __syncwarp();
            if (4 * lane < rem) *reinterpret_cast<int4*>(strip + 4 * lane) = keep;
            off += nblk * 128;
            fill = rem;
            __syncwarp();
        }
    }

    // ends a group's segment: pads it to whole loop iterations of the evaluation and writes the strip's remainder out
    __device__ __forceinline__ int finish_group(int pad_entry) {
        const int padn = (-(off + fill)) & (kPadTo2f - 1);
        for (int e = lane; e < padn; e += 32) strip[fill + e] = pad_entry;
        fill += padn;
        __syncwarp();
        if (off + fill <= pair_cap)
            for (int e = lane; e < fill; e += 32) list[off + e] = strip[e];
        off += fill;
        fill = 0;
        __syncwarp();
        return padn;
    }

    // true when staged neighbours n1 and n2 sit at exactly the same position as seen from the centre (d_ij == d_ik in
    // every component: the reference's r_jk is then 0 and it drops the triplet, acsf.py:316-325)
    template <typename T>
    __device__ __forceinline__ bool coincident(const AtomArgs<T>& a, int n1, int n2) const {
        T lx, ly, lz;
        bool pbc;
        item_box(a, slot, lx, ly, lz, pbc);
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        const Rec<T> ri = a.rec[slot];
        const Rec<T> r1 = a.rec[row[n1]], r2 = a.rec[row[n2]];
        T d1[3] = {sub_rn(ri.x, r1.x), sub_rn(ri.y, r1.y), sub_rn(ri.z, r1.z)};
        T d2[3] = {sub_rn(ri.x, r2.x), sub_rn(ri.y, r2.y), sub_rn(ri.z, r2.z)};
        if (pbc) {
            d1[0] = min_image(d1[0], lx); d1[1] = min_image(d1[1], ly); d1[2] = min_image(d1[2], lz);
            d2[0] = min_image(d2[0], lx); d2[1] = min_image(d2[1], ly); d2[2] = min_image(d2[2], lz);
        }
        return d1[0] == d2[0] && d1[1] == d2[1] && d1[2] == d2[2];
    }

    // one column block: 8 swept neighbours against the lane's 32 resident rows; returns the 8 sign bits (MSB first)
    template <bool SCREEN>
    __device__ __forceinline__ unsigned column_block(float bv, float bmul, float k4, float k5, const float (&ra)[2][4]) const {
        unsigned mask = 0;
        float d[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) mma_tf32_16x8x8(d[h], ra[h], bv, k4);
        if (SCREEN) {  // second test: (r_jk^2 + r_s^2 + r_l^2 - r2max) / 2
            const float bv2 = bv * bmul;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float e[4];
                mma_tf32_16x8x8(e, ra[h], bv2, k5);
#pragma unroll
                for (int j = 0; j < 4; ++j) d[h][j] = fmaxf(d[h][j], e[j]);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j) mask = __funnelshift_l(__float_as_uint(d[h][j]), mask, 1);
        return mask;
    }

    // resident chunk (staged neighbours r0 .. r0 + nres - 1, nres <= 32; their record offset goes into the entry half
    // selected by res_shift) against the swept range [s_begin, s_begin + ns): tiles of 32 x 32 pair tests on the tensor cores
    // tri: resident chunk and swept range are the same neighbours -- only the pairs with the swept index below the
    // resident one are kept (lower triangle of the tile)
    template <typename T>
    __device__ __forceinline__ void rect(const AtomArgs<T>& a, int r0, int nres, int res_shift, int s_begin, int ns, bool tri) {
        const float* sf = reinterpret_cast<const float*>(sf4);
        const int g = lane >> 2, t = lane & 3;
        float ra[2][4];
        int rpart[4];
        const float h1 = 0.5f * thr1, h2 = 0.5f * r2max;
        const float* pr = sf + (r0 + 4 * g) * 4;  // the lane's four rows: residents r0 + 4 g .. + 3 (rows past nres: masked)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = 2 * h + u;
                const float v = pr[4 * i + t], sq = pr[4 * i + 3];
                ra[h][u] = t == 3 ? 1.f : -v;
                const float l1 = sq - h1, l2 = fmaf(2.f, sq, -h2);  // not rounded: the tensor core drops their low 13 bits (in the margin)
                ra[h][2 + u] = t == 0 ? l1 : (t == 1 ? l2 : 0.f);
            }
        const int rstep = rb << res_shift;
        rpart[0] = ((r0 + 4 * g) * rb) << res_shift;
#pragma unroll
        for (int i = 1; i < 4; ++i) rpart[i] = rpart[i - 1] + rstep;
        unsigned rowmask = kRowMask2.m[nres][g];
        if (cls_test) {  // rows reach beyond the cutoff (skin): residents past it do not count
            unsigned rowbyte = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (2.f * pr[4 * i + 3] < thr1) rowbyte |= 0xC0u >> (2 * i);
            rowmask &= rowbyte * 0x01010101u;
        }
        if (tri) rowmask &= kTriMask2.m[lane];
        const float bmul = t == 3 ? 2.f : 1.f, k4 = t == 0 ? 1.f : 0.f, k5 = t == 1 ? 1.f : 0.f;
        const int stride = rb << (16 - res_shift);
        const float* pb = sf + s_begin * 4 + t;  // a partial tile reads past the range (inside the staging array)
        for (int s0 = 0; s0 < ns; s0 += 32, pb += 128) {
            const int nb = min(32, ns - s0);
            const int ncb = (nb + 7) >> 3;
            unsigned mask = 0;
            if (nb == 32) {  // four independent blocks
                unsigned m4[4];
                if (screen) {
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) m4[cb] = column_block<true>(pb[32 * cb + 4 * ((g + 2 * cb) & 7)], bmul, k4, k5, ra);
                } else {
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) m4[cb] = column_block<false>(pb[32 * cb + 4 * ((g + 2 * cb) & 7)], bmul, k4, k5, ra);
                }
                mask = __byte_perm(__byte_perm(m4[3], m4[2], 0x0040), __byte_perm(m4[1], m4[0], 0x0040), 0x5410);
            } else {
#pragma unroll 1
                for (int cb = 0; cb < ncb; ++cb) {
                    const float bv = pb[32 * cb + 4 * ((g + 2 * cb) & 7)];
                    mask = (mask << 8) | (screen ? column_block<true>(bv, bmul, k4, k5, ra) : column_block<false>(bv, bmul, k4, k5, ra));
                }
                mask <<= 8 * (4 - ncb);
            }
            unsigned valid = rowmask & kColMask2.m[nb][t];
            if (cls_test) {  // with a skin: swept neighbours past the cutoff do not count
                unsigned colmask = 0;
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    const int col = tile_col(t, cb, 0);  // the lane's two columns of a block are neighbours col, col + 1
                    const bool ok0 = 2.f * sf[(s_begin + s0 + col) * 4 + 3] < thr1, ok1 = 2.f * sf[(s_begin + s0 + col + 1) * 4 + 3] < thr1;
                    colmask |= ((ok0 ? 0xAAu : 0u) | (ok1 ? 0x55u : 0u)) << (24 - 8 * cb);
                }
                valid &= colmask;
            }
            mask &= valid;
            if (PANTEA_EXACT_ON && exact) {  // rare: drop the pairs of exactly coincident neighbours (candidates: identical staged vectors)
                unsigned zz = mask;
                while (zz) {
                    const int bit = 31 - __clz(zz);
                    zz &= ~(1u << bit);
                    const int k = 31 - bit;
                    const int rw = r0 + tile_row(g, (k >> 2) & 1, (k >> 1) & 1);
                    const int cl = s_begin + s0 + tile_col(t, k >> 3, k & 1);
                    const float4 fs = sf4[cl], fl = sf4[rw];
                    if (fs.x == fl.x && fs.y == fl.y && fs.z == fl.z && coincident<T>(a, rw, cl)) mask &= ~(1u << bit);
                }
            }
            emit(mask, ncb, rpart, ((s_begin + s0) * rb) << (16 - res_shift), stride);
        }
    }
};

#ifndef PANTEA_FILTER2_MINBLOCKS
#define PANTEA_FILTER2_MINBLOCKS 7
#endif
template <typename T>
__device__ __forceinline__ void filter2_atom(const AtomArgs<T>& a, int w, int lane, int wib, unsigned char* smem_raw) {
    if (w >= a.n_work) return;
    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    int32_t* off_out = a.pair_off + (size_t)w * (a.max_groups + 1);
    if (etype >= a.n_types) {
        if (lane == 0) off_out[0] = 0;
        return;
    }
    const ElementTable& tab = a.tables[etype];
    float4* sf4 = (float4*)smem_raw + (size_t)wib * (a.scap + 32);  // + 32: a partial tile may read past the row
    int32_t* strip = (int32_t*)((float4*)smem_raw + (size_t)kFilter2Warps * (a.scap + 32)) + wib * kStrip2;

    Segments sg;
    sg.load(a.tcount + (size_t)slot * kBuckets, a.scap);
    if (sg.seg[kBuckets] > a.scap && lane == 0) atomicMax(&a.flags[1], sg.seg[kBuckets]);

    T lx, ly, lz;
    bool pbc;
    item_box(a, slot, lx, ly, lz, pbc);

    // stage (dx, dy, dz, r^2 / 2): differences formed in T, then rounded to TF32
    const float rcf = tab.n_cls > 0 ? (float)tab.cls[0].rc + a.skin : 0.f;
    const float rc2f = rcf * rcf * 1.0001f + 1e-4f;  // inclusive: the exact test is the evaluation's
    const float tf32_margin = 5.0e-3f * (float)(a.rc_list * a.rc_list) + 1e-3f;  // rounded vectors and constants (see mma_tf32_16x8x8)
    {
        const Rec<T> ri = a.rec[slot];
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        // all index loads of up to five 32-neighbour rounds in flight, then all record gathers
        constexpr int RB = 5;
        for (int base = 0; base < sg.total; base += 32 * RB) {
            int idx[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n = base + 32 * r + lane;
                idx[r] = row[n < sg.total ? n : sg.total - 1];
            }
            Rec<T> rr[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) rr[r] = a.rec[idx[r]];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n = base + 32 * r + lane;
                if (n < sg.total) {
                    T dx = sub_rn(ri.x, rr[r].x), dy = sub_rn(ri.y, rr[r].y), dz = sub_rn(ri.z, rr[r].z);
                    if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
                    const float fx = to_tf32((float)dx), fy = to_tf32((float)dy), fz = to_tf32((float)dz);
                    sf4[n] = make_float4(fx, fy, fz, to_tf32(0.5f * (fx * fx + fy * fy + fz * fz)));
                }
            }
        }
    }
    __syncwarp();

    Filter2 f;
    f.sf4 = sf4; f.list = a.pairs + (size_t)w * a.pair_cap; f.strip = strip; f.pair_cap = a.pair_cap; f.off = 0; f.fill = 0;
    f.lane = lane; f.thr1 = rc2f + tf32_margin;
    f.slot = slot; f.exact = a.dup_flag && *a.dup_flag != 0;
    f.cls_test = tab.n_cls > 0 && tab.cls[0].rc + (double)a.skin < a.rc_list;  // rows reach beyond the cutoff
    f.rb = a.rec_bytes;
    const int pad_entry = (sg.total * f.rb) | ((sg.total * f.rb) << 16);
    int n_real = 0;
    for (int gi = 0; gi < tab.n_groups; ++gi) {
        if (lane == 0) off_out[gi] = f.off < f.pair_cap ? f.off : f.pair_cap;
        const int seg_begin = f.off;
        const AngularGroup grp = tab.groups[gi];
        const int bj = sg.lo(grp.type_j), nj = sg.hi(grp.type_j) - bj;
        const int bk = sg.lo(grp.type_k), nk = sg.hi(grp.type_k) - bk;
        // Entries hold the type_j member's record offset in the low half and the type_k member's in the high half
        // (same-type groups: either way).  Full 32-lane chunks of one bucket stay resident while the other bucket is
        // swept; the chunk remainder is swept against resident chunks of the other bucket instead (few idle lanes
        // either way); the order of the pairs is fixed, hence so is the evaluation's summation order.
        // Gaussian screening (pantea_set_gauss_screen): reference pair = nearest neighbour of type_j with nearest of type_k
        // (same type: the two nearest); it is a live triplet when the two are closer than the cutoff to each other, which
        // r_j + r_k < rc guarantees.  Pairs whose weight is below exp(-T) of its weight cannot matter and are dropped.
        f.screen = false;
        f.r2max = 3.0e38f;
        if (a.screen_t > 0.f && nj > 0 && nk > 0) {
            const float eta = (float)tab.members[grp.first].eta;
            const float reach = a.screen_t / fmaxf(eta, 1e-30f);  // T / eta
            if (reach < 3.f * rc2f) {  // otherwise nothing inside the cutoff sphere can be dropped
                float b1 = 3.0e38f, b2 = 3.0e38f;  // smallest r^2 of the j bucket / of the k bucket (same type: second smallest)
                int i1 = -1, i2 = -1;
                for (int n = lane; n < nj; n += 32) { const float v = 2.f * sf4[bj + n].w; if (v < b1) { b1 = v; i1 = bj + n; } }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float v = __shfl_xor_sync(kFullMask, b1, o); const int iv = __shfl_xor_sync(kFullMask, i1, o);
                    if (v < b1 || (v == b1 && iv < i1)) { b1 = v; i1 = iv; }
                }
                for (int n = lane; n < nk; n += 32) { const float v = 2.f * sf4[bk + n].w; if (v < b2 && bk + n != i1) { b2 = v; i2 = bk + n; } }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float v = __shfl_xor_sync(kFullMask, b2, o); const int iv = __shfl_xor_sync(kFullMask, i2, o);
                    if (v < b2 || (v == b2 && iv < i2)) { b2 = v; i2 = iv; }
                }
                if (i1 >= 0 && i2 >= 0 && sqrtf(b1) + sqrtf(b2) < 0.999f * ((float)tab.cls[grp.cls].rc)) {
                    const float4 p1 = sf4[i1], p2 = sf4[i2];
                    const float ex = p1.x - p2.x, ey = p1.y - p2.y, ez = p1.z - p2.z;
                    const float ref = b1 + b2 + ex * ex + ey * ey + ez * ez;
                    f.r2max = (ref + reach) * 1.001f + 2.f * tf32_margin;  // the three squared distances carry TF32 rounding
                    f.screen = true;
                }
            }
        }
        // The work is enumerated as items (resident chunk, swept range, diagonal flag) so that the tile and emission code
        // is instantiated once (instruction cache).
        const bool same = grp.type_j == grp.type_k;
        auto cost = [](int nr, int ns) { return (nr >> 5) * ns + ((nr & 31) ? ((ns + 31) >> 5) * (nr & 31) : 0); };
        const bool res_k = same || cost(nk, nj) <= cost(nj, nk);
        const int rb = res_k ? bk : bj, nr = res_k ? nk : nj;  // bucket whose full chunks stay resident
        const int sb = res_k ? bj : bk, ns = res_k ? nj : nk;  // (same-type: both are the one bucket)
        const int rshift = res_k ? 16 : 0;
        const int full = nr >> 5, rem = nr & 31;
        const int n_tail = rem == 0 ? 0 : (same ? full + 1 : (ns + 31) >> 5);
        const int n_items = (full + n_tail) << (same ? 1 : 0);  // same-type groups: odd items are the chunks' own triangles
        for (int it2 = 0; it2 < n_items; ++it2) {
            const int it = same ? it2 >> 1 : it2;
            int r0, nres, shift, s_begin, s_len;
            if (it < full) {  // a full resident chunk: (same-type) everything before it, else everything
                r0 = rb + 32 * it; nres = 32; shift = rshift; s_begin = sb; s_len = same ? 32 * it : ns;
            } else if (same) {  // remainder of the bucket: swept against the full chunks
                const int c = it - full;
                if (c < full) { r0 = rb + 32 * c; nres = 32; shift = 16; s_begin = rb + 32 * full; s_len = rem; }
                else { r0 = rb + 32 * full; nres = rem; shift = 16; s_begin = 0; s_len = 0; }
            } else {  // remainder of the resident bucket: swept against resident chunks of the other bucket
                const int q = it - full;
                r0 = sb + 32 * q; nres = min(32, ns - 32 * q); shift = 16 - rshift; s_begin = rb + 32 * full; s_len = rem;
            }
            bool tri = false;
            if (same && (it2 & 1)) {  // the chunk's own triangle: a masked tile
                const bool own = it < full || it == 2 * full;  // full chunks and the remainder chunk (not the remainder-vs-chunk items)
                tri = true; shift = 16; s_begin = r0; s_len = (own && nres > 1) ? nres : 0;
            }
            if (s_len > 0) f.template rect<T>(a, r0, nres, shift, s_begin, s_len, tri);
        }
        n_real -= f.finish_group(pad_entry);
        n_real += f.off - seg_begin;
    }
    if (lane == 0) {
        off_out[tab.n_groups] = f.off < f.pair_cap ? f.off : f.pair_cap;
        atomicMax(&a.flags[2], f.off);
        if (a.counters) atomicAdd(&a.counters[3], (unsigned long long)n_real);  // list entries the evaluation will walk
    }
    __syncwarp();  // the warp's staging array and strip are reused by its next atom
}

template <typename T>
__global__ void __launch_bounds__(kFilter2Warps * 32, PANTEA_FILTER2_MINBLOCKS) pair_filter2_kernel(const AtomArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.filter_guard && *a.filter_guard == 0) return;  // rows unchanged since the lists were written
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (kFilterAtomsPerWarp == 1) {  // one atom per warp, no counter
        filter2_atom<T>(a, blockIdx.x * kFilter2Warps + wib, lane, wib, smem_raw);
        return;
    }
    __shared__ int s_next;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    constexpr int per_block = kFilter2Warps * kFilterAtomsPerWarp;
    for (int k = next_item(&s_next, lane); k < per_block; k = next_item(&s_next, lane))
        filter2_atom<T>(a, blockIdx.x * per_block + k, lane, wib, smem_raw);
}

// ------------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------------
// One angular group over its padded pair list.  ZM1 = zeta - 1 as a compile-time constant (0, 1, 3) or -1: runtime.
#ifndef PANTEA_NU2
#define PANTEA_NU2 2  // triplets per lane and loop iteration: independent dependency chains for the in-order issue
#endif
#ifndef PANTEA_EVAL2_MINBLOCKS
#define PANTEA_EVAL2_MINBLOCKS 4
#endif
constexpr int kNU2 = PANTEA_NU2;

template <int ZM1>
__device__ __forceinline__ void angular2(const unsigned char* __restrict__ snb, const int32_t* __restrict__ list, int n_iter,
                                         double neta, double lam, double zl, int zm1_rt, double rc, int lane, int* stage,
                                         const double* __restrict__ etab, bool first_staged, double& oG, double& oX, double& oY,
                                         double& oZ) {
    constexpr int NU = kNU2;
    const double m2_inv_rc = -2.0 / rc;
    double aG[NU], aX[NU], aY[NU], aZ[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) { aG[u] = 0; aX[u] = 0; aY[u] = 0; aZ[u] = 0; }
    constexpr int CH = kChunk2;  // 32-entry rows per cp.async group (a multiple of 4 and of NU)
    const int n_chunks = (n_iter + CH - 1) / CH;
    // every lane copies 16 bytes per 4 rows (4 x 32 entries = 512 bytes) of the chunk into the ring
    auto stage_chunk = [&](int chunk) {
        int* dst = stage + (chunk & 1) * (CH * 32);
        const int32_t* src = list + (size_t)chunk * (CH * 32);
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) cp_async16(dst + 128 * q + 4 * lane, src + 128 * q + 4 * lane);
        cp_async_commit();
    };
    if (n_chunks > 0 && !first_staged) stage_chunk(0);  // first_staged: the caller issued chunk 0 before its record arithmetic
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        if (chunk + 1 < n_chunks) {
            stage_chunk(chunk + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();  // the chunk was copied by all lanes of the warp
        const int* src = stage + (chunk & 1) * (CH * 32) + lane;
        const int n_in = min(CH, n_iter - chunk * CH);
        for (int i = 0; i < n_in; i += NU, src += 32 * NU) {
            D2 j0[NU], j1[NU], j2[NU], j3[NU], k0[NU], k1[NU], k2[NU], k3[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int jk = src[32 * u];
                const unsigned char* pj = snb + (jk & 0xffff);
                const unsigned char* pk = snb + ((unsigned)jk >> 16);
                // [u_x u_y | u_z r^2 | 1/r sqrt2 r | W Q]
                j0[u] = lds128(pj); j1[u] = lds128(pj + 16); j2[u] = lds128(pj + 32); j3[u] = lds128(pj + 48);
                k0[u] = lds128(pk); k1[u] = lds128(pk + 16); k2[u] = lds128(pk + 32); k3[u] = lds128(pk + 48);
            }
            double cost[NU], rjk2[NU], rjk[NU], x2[NU], th[NU], g[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                cost[u] = fma(j1[u].a, k1[u].a, fma(j0[u].b, k0[u].b, j0[u].a * k0[u].a));
                // r_jk^2 = r_j^2 + r_k^2 - 2 r_j r_k cos, kept above 1e-30 (rounding can leave zero or a negative
                // residual for neighbours closer than 1e-7 Bohr to each other; exactly coincident ones never reach the list)
                rjk2[u] = clamp_hi(fma(-(j2[u].b * k2[u].b), cost[u], j1[u].b + k1[u].b), 0x39b00000);
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) rjk[u] = sqrt_cubic(rjk2[u]);
            // tanh(x) = 1 - 2 / (exp(2x) + 1) with 2x = 2 - 2 r_jk / rc clamped at 0: exactly 0 at and beyond the cutoff
#pragma unroll
            for (int u = 0; u < NU; ++u) x2[u] = clamp_hi(fma(rjk[u], m2_inv_rc, 2.0), 0);
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                th[u] = exp_tab256(x2[u], etab) + 1.0;
                g[u] = exp_tab256(neta * rjk2[u], etab);
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) th[u] = fma(-2.0, rcp_cubic(th[u]), 1.0);
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const double bs = fma(lam, cost[u], 1.0);
                double ep = (j3[u].a * k3[u].a) * ((th[u] * th[u]) * (th[u] * g[u]));
                if (ZM1 == 1) ep *= bs;
                else if (ZM1 == 3) ep *= bs * (bs * bs);
                else if (ZM1 < 0) ep *= powi<double>(bs, zm1_rt);
                const double ap = bs * ep;
                aG[u] += ap;
                const double Tc = zl * ep;
                const double Bj = fma(Tc, fma(-cost[u], j2[u].a, k2[u].a), ap * j3[u].b);
                const double Bk = fma(Tc, fma(-cost[u], k2[u].a, j2[u].a), ap * k3[u].b);
                aX[u] = fma(Bk, k0[u].a, fma(Bj, j0[u].a, aX[u]));
                aY[u] = fma(Bk, k0[u].b, fma(Bj, j0[u].b, aY[u]));
                aZ[u] = fma(Bk, k1[u].a, fma(Bj, j1[u].a, aZ[u]));
            }
        }
        __syncwarp();  // all lanes are done with this half of the ring before it is refilled
    }
#pragma unroll
    for (int u = 1; u < NU; ++u) { aG[0] += aG[u]; aX[0] += aX[u]; aY[0] += aY[u]; aZ[0] += aZ[u]; }
    oG = warp_sum(aG[0]); oX = warp_sum(aX[0]); oY = warp_sum(aY[0]); oZ = warp_sum(aZ[0]);
}

__device__ __forceinline__ void eval2_atom(const AtomArgs<double>& a, int w, int lane, int wib, unsigned char* smem_raw,
                                           const double* __restrict__ s_etab, int* stage) {
    if (w >= a.n_work) return;
    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    if (etype >= a.n_types) return;
    const ElementTable& tab = a.tables[etype];
    const int n_sf = tab.n_sf;
    const Rec<double> ri = a.rec[slot];
    double lx, ly, lz;
    bool pbc;
    item_box(a, slot, lx, ly, lz, pbc);

    const int cap = a.scap;
    const size_t per_atom = eval2_smem_bytes(cap, a.n_sf_max);
    unsigned char* snb = smem_raw + (size_t)wib * per_atom;                    // [cap + 1] records of 80 bytes
    double* sacc = (double*)(snb + (size_t)(cap + 1) * kRec2Bytes);           // [n_sf_max][4]

    Segments sg;
    sg.load(a.tcount + (size_t)slot * kBuckets, cap);
    const int total = sg.total;
    if (sg.seg[kBuckets] > cap && lane == 0) atomicMax(&a.flags[1], sg.seg[kBuckets]);
    const int32_t* offs = a.pair_off + (size_t)w * (a.max_groups + 1);
    const int32_t* lists = a.pairs + (size_t)w * a.pair_cap;
    const int off0 = tab.n_groups > 0 ? offs[0] : 0, off1 = tab.n_groups > 0 ? offs[1] : 0;  // in flight behind the index loads

    // ---- stage the neighbour block --------------------------------------------------------------------------------
    {
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        const bool has_cls = tab.n_cls > 0;
        const int ct = has_cls ? tab.cls[0].type : PANTEA_CUT_HARD;
        const double rcc = has_cls ? tab.cls[0].rc : 1.0;
        const double irc = rcp_cubic(rcc);
        // phase 1 (memory): all index loads of up to five 32-neighbour rounds in flight, then all record gathers --
        // two dependent global latencies per atom instead of two per round; the difference vectors are parked in the
        // record slots
        constexpr int RB = 5;
        for (int base = 0; base < total; base += 32 * RB) {
            int idx[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n = base + 32 * r + lane;
                idx[r] = row[n < total ? n : total - 1];
            }
            Rec<double> rr[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) rr[r] = a.rec[idx[r]];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n = base + 32 * r + lane;
                if (n < total) {
                    double dx = sub_rn(ri.x, rr[r].x), dy = sub_rn(ri.y, rr[r].y), dz = sub_rn(ri.z, rr[r].z);
                    if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
                    double* p = (double*)(snb + (size_t)n * kRec2Bytes);
                    p[0] = dx; p[1] = dy; p[2] = dz; p[3] = __hiloint2double(0, rr[r].type);
                }
            }
        }
        // the first chunk of the first angular group's pair list streams into the ring while the records are computed
        if (off1 - off0 >= 32) {
            const int32_t* src = lists + off0;
#pragma unroll
            for (int q = 0; q < kChunk2 / 4; ++q) cp_async16(stage + 128 * q + 4 * lane, src + 128 * q + 4 * lane);
            cp_async_commit();
        }
        // phase 2 (arithmetic): every lane turns its own parked vectors into records, two at a time (independent
        // dependency chains: reciprocal root, two exponentials, two reciprocals each)
        auto make_record = [&](double dx, double dy, double dz, int ntype, double (&o)[kRec2]) {
            const double r2n = dx * dx + dy * dy + dz * dz;
            const double iv = fast_rsqrt(r2n), r = r2n * iv;
            double fc, q;
            if (ct == PANTEA_CUT_TANHU) {  // tanh(x) = 1 - 2 / (exp(2x) + 1), x = 1 - r / rc
                const double t = fma(-2.0, rcp_cubic(exp_tab256(fmax(fma(-2.0 * irc, r, 2.0), 0.0), s_etab) + 1.0), 1.0);
                const bool in = r < rcc && t > 0.0;
                fc = in ? t * t * t : 0.0;
                q = in ? -3.0 * irc * (1.0 - t * t) * rcp_cubic(t) : 0.0;
            } else {
                double dfc;
                cutoff_eval_ool<double>(ct, r, rcc, &fc, &dfc);
                q = fc != 0.0 ? dfc / fc : 0.0;
            }
            // W = sqrt(pref) fc exp(-eta r^2), Q = fc'/fc - 2 eta r for the group this neighbour type takes part in
            const double eta_t = tab.v2_eta[ntype], ws_t = tab.v2_wscale[ntype];
            const double W = ws_t * fc * exp_tab256(fmax(-eta_t * r2n, -700.0), s_etab);
            const double Q = fma(-2.0 * eta_t, r, q);
            o[0] = dx * iv; o[1] = dy * iv; o[2] = dz * iv; o[3] = r2n; o[4] = iv; o[5] = 1.4142135623730951 * r;
            o[6] = W; o[7] = Q; o[8] = fc; o[9] = q;
        };
        for (int n = lane; n < total; n += 64) {
            const bool two = n + 32 < total;
            double* pa = (double*)(snb + (size_t)n * kRec2Bytes);
            double* pb = (double*)(snb + (size_t)(two ? n + 32 : n) * kRec2Bytes);
            const double ax = pa[0], ay = pa[1], az = pa[2], bx = pb[0], by = pb[1], bz = pb[2];
            const int ta = __double2loint(pa[3]), tb = __double2loint(pb[3]);
            double ra[kRec2], rb[kRec2];
            make_record(ax, ay, az, ta, ra);
            make_record(bx, by, bz, tb, rb);
#pragma unroll
            for (int c = 0; c < kRec2; ++c) pa[c] = ra[c];
            if (two) {
#pragma unroll
                for (int c = 0; c < kRec2; ++c) pb[c] = rb[c];
            }
        }
        // the record pad entries point at: W = 0 makes every term vanish, r^2 = 1 keeps r_jk finite
        if (lane < kRec2) ((double*)(snb + (size_t)total * kRec2Bytes))[lane] = lane == 3 ? 1.0 : 0.0;
    }
    __syncwarp();

    // ---- radial symmetry functions (lanes over neighbours) ---------------------------------------------------------
    for (int s = 0; s < tab.n_radial; ++s) {
        const RadialSF sf = tab.radial[s];
        const int lo = sg.lo(sf.type_j), hi = sg.hi(sf.type_j);
        const double eta = sf.eta, rs = sf.r_shift;
        double g = 0, gx = 0, gy = 0, gz = 0;
        for (int n = lo + lane; n < hi; n += 32) {
            const double* p = (const double*)(snb + (size_t)n * kRec2Bytes);
            const double r = p[3] * p[4], fc = p[8], q = p[9];
            double val, dval;
            if (sf.kind == PANTEA_G1) { val = fc; dval = fc * q; }
            else {
                const double dr = r - rs, ex = fast_exp(-eta * dr * dr);
                val = ex * fc; dval = val * (q - 2.0 * eta * dr);
            }
            g += val;
            gx += dval * p[0]; gy += dval * p[1]; gz += dval * p[2];
        }
        g = warp_sum(g); gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
        if (lane == 0) { double* o = sacc + 4 * sf.out; o[0] = g; o[1] = gx; o[2] = gy; o[3] = gz; }
    }

    // ---- angular symmetry functions: flat walk over the padded pair lists -------------------------------------------
    {
        for (int gi = 0; gi < tab.n_groups; ++gi) {
            const AngularGroup grp = tab.groups[gi];
            const AngularMember mem = tab.members[grp.first];
            const int lo = gi == 0 ? off0 : offs[gi], n_iter = ((gi == 0 ? off1 : offs[gi + 1]) - lo) >> 5;
            const bool pre = gi == 0 && off1 - off0 >= 32;  // chunk 0 is already on its way
            const double rc = tab.cls[grp.cls].rc;
            const double zl = mem.zeta * mem.lambda0;  // pref lives in the W weights
            double G, X, Y, Z;
            if (mem.izeta == 1) angular2<0>(snb, lists + lo, n_iter, -mem.eta, mem.lambda0, zl, 0, rc, lane, stage, s_etab, pre, G, X, Y, Z);
            else if (mem.izeta == 2) angular2<1>(snb, lists + lo, n_iter, -mem.eta, mem.lambda0, zl, 1, rc, lane, stage, s_etab, pre, G, X, Y, Z);
            else if (mem.izeta == 4) angular2<3>(snb, lists + lo, n_iter, -mem.eta, mem.lambda0, zl, 3, rc, lane, stage, s_etab, pre, G, X, Y, Z);
            else angular2<-1>(snb, lists + lo, n_iter, -mem.eta, mem.lambda0, zl, mem.izeta - 1, rc, lane, stage, s_etab, pre, G, X, Y, Z);
            if (lane == 0) { double* o = sacc + 4 * mem.out; o[0] = G; o[1] = X; o[2] = Y; o[3] = Z; }
        }
    }
    __syncwarp();
    if (a.G)
        for (int s = lane; s < n_sf; s += 32) a.G[(size_t)out_row * a.g_stride + s] = sacc[4 * s];
    if (a.dG)
        for (int e = lane; e < n_sf * 3; e += 32) {
            const int s = e / 3, c = e - 3 * s;
            a.dG[((size_t)out_row * a.g_stride + s) * 3 + c] = sacc[4 * s + 1 + c];
        }
    if (a.gbuf)
        for (int e = lane; e < n_sf * 4; e += 32) a.gbuf[(size_t)w * a.n_sf_max * 4 + e] = sacc[e];
    __syncwarp();  // the warp's shared-memory regions are reused by its next atom
}

__global__ void __launch_bounds__(kEvalWarps * 32, PANTEA_EVAL2_MINBLOCKS) hdnnp_eval2_kernel(const AtomArgs<double> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    __shared__ double s_etab[kTab2];
    for (int i = threadIdx.x; i < kTab2; i += blockDim.x) s_etab[i] = exp2((double)i * (1.0 / kTab2));
    __shared__ __align__(16) int s_stage[kEvalWarps][2 * kChunk2 * 32];
    __shared__ int s_next;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int per_block = kEvalWarps * a.apw;
    for (int k = next_item(&s_next, lane); k < per_block; k = next_item(&s_next, lane))
        eval2_atom(a, blockIdx.x * per_block + k, lane, wib, smem_raw, s_etab, s_stage[wib]);
}

// ------------------------------------------------------------------------------------------------
// evaluation in single precision (mixed mode of a double-precision workspace: pantea_workspace_set_compute_precision)
// ------------------------------------------------------------------------------------------------
// Positions, velocities and forces stay double; the difference vectors d_ij are formed in double (exact minimum image,
// no loss at large box lengths) and only then rounded, and everything from there to the summed symmetry functions is
// FP32: 48-byte records [d_x d_y d_z r^2 | 1/r W Q - | fc q - -] (12-bank stride: conflict-free LDS.128), r_jk from the
// difference of the two vectors (the dot-product form cancels too much in FP32), MUFU reciprocal root / exp2 /
// reciprocal.  The sums leave the kernel as doubles; scaler and network stay double (mlp_force_kernel).  Tolerance of
// the mode: 1e-5 relative (north star).
constexpr int kRec2fBytes = 48;
#ifndef PANTEA_EVAL2F_MINBLOCKS
#define PANTEA_EVAL2F_MINBLOCKS 6
#endif

__host__ __device__ inline size_t eval2f_smem_bytes(int scap, int n_sf) {
    return ((size_t)(scap + 1) * kRec2fBytes + (size_t)n_sf * 4 * 8 + 15) & ~size_t(15);
}

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqf(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float4 lds128f(const unsigned char* p) { return *reinterpret_cast<const float4*>(p); }

template <int ZM1>
__device__ __forceinline__ void angular2f(const unsigned char* __restrict__ snb, const int32_t* __restrict__ list, int n_iter,
                                          float neta_l2e, float lam, float zl, int zm1_rt, float rc, int lane, int* stage,
                                          bool first_staged, double& oG, double& oX, double& oY, double& oZ) {
    constexpr int NU = kNU2;
    constexpr float kL2E = 1.4426950408889634f;
    const float m2l_inv_rc = -2.0f * kL2E / rc, two_l2e = 2.0f * kL2E;
    float aG[NU], aX[NU], aY[NU], aZ[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) { aG[u] = 0; aX[u] = 0; aY[u] = 0; aZ[u] = 0; }
    constexpr int CH = kChunk2;
    const int n_chunks = (n_iter + CH - 1) / CH;
    auto stage_chunk = [&](int chunk) {
        int* dst = stage + (chunk & 1) * (CH * 32);
        const int32_t* src = list + (size_t)chunk * (CH * 32);
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) cp_async16(dst + 128 * q + 4 * lane, src + 128 * q + 4 * lane);
        cp_async_commit();
    };
    if (n_chunks > 0 && !first_staged) stage_chunk(0);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        if (chunk + 1 < n_chunks) {
            stage_chunk(chunk + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const int* src = stage + (chunk & 1) * (CH * 32) + lane;
        const int n_in = min(CH, n_iter - chunk * CH);
        for (int i = 0; i < n_in; i += NU, src += 32 * NU) {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int jk = src[32 * u];
                const unsigned char* pj = snb + (jk & 0xffff);
                const unsigned char* pk = snb + ((unsigned)jk >> 16);
                const float4 j0 = lds128f(pj), j1 = lds128f(pj + 16);  // [d_x d_y d_z r^2 | 1/r W Q -]
                const float4 k0 = lds128f(pk), k1 = lds128f(pk + 16);
                const float ex = j0.x - k0.x, ey = j0.y - k0.y, ez = j0.z - k0.z;
                const float rjk2 = fmaxf(fmaf(ez, ez, fmaf(ey, ey, ex * ex)), 1e-30f);
                const float rjk = rjk2 * rsqf(rjk2);
                const float cost = fmaf(j0.z, k0.z, fmaf(j0.y, k0.y, j0.x * k0.x)) * (j1.x * k1.x);
                // tanh(x) = 1 - 2 / (exp(2x) + 1), 2x = 2 - 2 r_jk / rc clamped at 0: exactly 0 at and beyond the cutoff
                const float x2 = fmaxf(fmaf(rjk, m2l_inv_rc, two_l2e), 0.0f);
                const float th = fmaf(-2.0f, rcpf(ex2f(x2) + 1.0f), 1.0f);
                const float g = ex2f(neta_l2e * rjk2);
                const float bs = fmaf(lam, cost, 1.0f);
                float ep = (j1.y * k1.y) * ((th * th) * (th * g));
                if (ZM1 == 1) ep *= bs;
                else if (ZM1 == 3) ep *= bs * (bs * bs);
                else if (ZM1 < 0) ep *= powi<float>(bs, zm1_rt);
                const float ap = bs * ep;
                aG[u] += ap;
                const float Tc = zl * ep;
                // gradient coefficients on the difference vectors: (B / r) d
                const float Bj = fmaf(Tc, fmaf(-cost, j1.x, k1.x), ap * j1.z) * j1.x;
                const float Bk = fmaf(Tc, fmaf(-cost, k1.x, j1.x), ap * k1.z) * k1.x;
                aX[u] = fmaf(Bk, k0.x, fmaf(Bj, j0.x, aX[u]));
                aY[u] = fmaf(Bk, k0.y, fmaf(Bj, j0.y, aY[u]));
                aZ[u] = fmaf(Bk, k0.z, fmaf(Bj, j0.z, aZ[u]));
            }
        }
        __syncwarp();
    }
    double dG = 0, dX = 0, dY = 0, dZ = 0;
#pragma unroll
    for (int u = 0; u < NU; ++u) { dG += (double)aG[u]; dX += (double)aX[u]; dY += (double)aY[u]; dZ += (double)aZ[u]; }
    oG = warp_sum(dG); oX = warp_sum(dX); oY = warp_sum(dY); oZ = warp_sum(dZ);
}

__device__ __forceinline__ void eval2f_atom(const AtomArgs<double>& a, int w, int lane, int wib, unsigned char* smem_raw, int* stage) {
    if (w >= a.n_work) return;
    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    if (etype >= a.n_types) return;
    const ElementTable& tab = a.tables[etype];
    const int n_sf = tab.n_sf;
    const Rec<double> ri = a.rec[slot];
    double lx, ly, lz;
    bool pbc;
    item_box(a, slot, lx, ly, lz, pbc);
    constexpr float kL2E = 1.4426950408889634f;

    const int cap = a.scap;
    const size_t per_atom = eval2f_smem_bytes(cap, a.n_sf_max);
    unsigned char* snb = smem_raw + (size_t)wib * per_atom;                     // [cap + 1] records of 48 bytes
    double* sacc = (double*)(snb + (((size_t)(cap + 1) * kRec2fBytes + 7) & ~size_t(7)));  // [n_sf_max][4]

    Segments sg;
    sg.load(a.tcount + (size_t)slot * kBuckets, cap);
    const int total = sg.total;
    if (sg.seg[kBuckets] > cap && lane == 0) atomicMax(&a.flags[1], sg.seg[kBuckets]);

    // ---- stage the neighbour block: gathers and differences in double, records in single -------------------------
    {
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        const bool has_cls = tab.n_cls > 0;
        const int ct = has_cls ? tab.cls[0].type : PANTEA_CUT_HARD;
        const float rcc = has_cls ? (float)tab.cls[0].rc : 1.0f;
        const float irc = 1.0f / rcc;
        constexpr int RB = 5;
        for (int base = 0; base < total; base += 32 * RB) {
            int idx[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n = base + 32 * r + lane;
                idx[r] = row[n < total ? n : total - 1];
            }
            Rec<double> rr[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) rr[r] = a.rec[idx[r]];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int n = base + 32 * r + lane;
                if (n < total) {
                    double ddx = sub_rn(ri.x, rr[r].x), ddy = sub_rn(ri.y, rr[r].y), ddz = sub_rn(ri.z, rr[r].z);
                    if (pbc) { ddx = min_image(ddx, lx); ddy = min_image(ddy, ly); ddz = min_image(ddz, lz); }
                    const float dx = (float)ddx, dy = (float)ddy, dz = (float)ddz;
                    const int ntype = rr[r].type;
                    const float r2n = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float iv = rsqf(r2n), rad = r2n * iv;
                    float fc, q;
                    if (ct == PANTEA_CUT_TANHU) {
                        const float t = fmaf(-2.0f, rcpf(ex2f(fmaxf(fmaf(-2.0f * kL2E * irc, rad, 2.0f * kL2E), 0.0f)) + 1.0f), 1.0f);
                        const bool in = rad < rcc && t > 0.0f;
                        fc = in ? t * t * t : 0.0f;
                        q = in ? -3.0f * irc * (1.0f - t * t) * rcpf(t) : 0.0f;
                    } else {
                        float dfc;
                        cutoff_eval_ool<float>(ct, rad, rcc, &fc, &dfc);
                        q = fc != 0.0f ? dfc / fc : 0.0f;
                    }
                    const float eta_t = (float)tab.v2_eta[ntype], ws_t = (float)tab.v2_wscale[ntype];
                    const float W = ws_t * fc * ex2f(fmaxf(-eta_t * kL2E * r2n, -120.0f));
                    const float Q = fmaf(-2.0f * eta_t, rad, q);
                    float4* p = (float4*)(snb + (size_t)n * kRec2fBytes);
                    p[0] = make_float4(dx, dy, dz, r2n);
                    p[1] = make_float4(iv, W, Q, 0.0f);
                    p[2] = make_float4(fc, q, 0.0f, 0.0f);
                }
            }
        }
        // the record pad entries point at: W = 0 makes every term vanish, d = (1, 0, 0) keeps r_jk and 1/r finite
        if (lane < 3) ((float4*)(snb + (size_t)total * kRec2fBytes))[lane] = lane == 0 ? make_float4(0.f, 0.f, 0.f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    // ---- radial symmetry functions -----------------------------------------------------------------------------------
    for (int s = 0; s < tab.n_radial; ++s) {
        const RadialSF sf = tab.radial[s];
        const int lo = sg.lo(sf.type_j), hi = sg.hi(sf.type_j);
        const float eta = (float)sf.eta, rs = (float)sf.r_shift;
        float g = 0, gx = 0, gy = 0, gz = 0;
        for (int n = lo + lane; n < hi; n += 32) {
            const float4* p = (const float4*)(snb + (size_t)n * kRec2fBytes);
            const float4 p0 = p[0], p1 = p[1], p2 = p[2];
            const float r = p0.w * p1.x, fc = p2.x, q = p2.y;
            float val, dval;
            if (sf.kind == PANTEA_G1) { val = fc; dval = fc * q; }
            else {
                const float dr = r - rs, ex = ex2f(fmaxf(-eta * kL2E * dr * dr, -120.0f));
                val = ex * fc; dval = val * (q - 2.0f * eta * dr);
            }
            g += val;
            dval *= p1.x;  // on the difference vector: (dval / r) d
            gx += dval * p0.x; gy += dval * p0.y; gz += dval * p0.z;
        }
        const double G = warp_sum((double)g), X = warp_sum((double)gx), Y = warp_sum((double)gy), Z = warp_sum((double)gz);
        if (lane == 0) { double* o = sacc + 4 * sf.out; o[0] = G; o[1] = X; o[2] = Y; o[3] = Z; }
    }

    // ---- angular symmetry functions ----------------------------------------------------------------------------------
    {
        const int32_t* offs = a.pair_off + (size_t)w * (a.max_groups + 1);
        const int32_t* lists = a.pairs + (size_t)w * a.pair_cap;
        for (int gi = 0; gi < tab.n_groups; ++gi) {
            const AngularGroup grp = tab.groups[gi];
            const AngularMember mem = tab.members[grp.first];
            const int lo = offs[gi], n_iter = (offs[gi + 1] - lo) >> 5;
            const float rc = (float)tab.cls[grp.cls].rc;
            const float zl = (float)(mem.zeta * mem.lambda0), neta = (float)(-mem.eta) * kL2E, lam = (float)mem.lambda0;
            double G, X, Y, Z;
            if (mem.izeta == 1) angular2f<0>(snb, lists + lo, n_iter, neta, lam, zl, 0, rc, lane, stage, false, G, X, Y, Z);
            else if (mem.izeta == 2) angular2f<1>(snb, lists + lo, n_iter, neta, lam, zl, 1, rc, lane, stage, false, G, X, Y, Z);
            else if (mem.izeta == 4) angular2f<3>(snb, lists + lo, n_iter, neta, lam, zl, 3, rc, lane, stage, false, G, X, Y, Z);
            else angular2f<-1>(snb, lists + lo, n_iter, neta, lam, zl, mem.izeta - 1, rc, lane, stage, false, G, X, Y, Z);
            if (lane == 0) { double* o = sacc + 4 * mem.out; o[0] = G; o[1] = X; o[2] = Y; o[3] = Z; }
        }
    }
    __syncwarp();
    if (a.G)
        for (int s = lane; s < n_sf; s += 32) a.G[(size_t)out_row * a.g_stride + s] = sacc[4 * s];
    if (a.dG)
        for (int e = lane; e < n_sf * 3; e += 32) {
            const int s = e / 3, c = e - 3 * s;
            a.dG[((size_t)out_row * a.g_stride + s) * 3 + c] = sacc[4 * s + 1 + c];
        }
    if (a.gbuf)
        for (int e = lane; e < n_sf * 4; e += 32) a.gbuf[(size_t)w * a.n_sf_max * 4 + e] = sacc[e];
    __syncwarp();
}

__global__ void __launch_bounds__(kEvalWarps * 32, PANTEA_EVAL2F_MINBLOCKS) hdnnp_eval2f_kernel(const AtomArgs<double> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    __shared__ __align__(16) int s_stage[kEvalWarps][2 * kChunk2 * 32];
    __shared__ int s_next;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int per_block = kEvalWarps * a.apw;
    for (int k = next_item(&s_next, lane); k < per_block; k = next_item(&s_next, lane))
        eval2f_atom(a, blockIdx.x * per_block + k, lane, wib, smem_raw, s_stage[wib]);
}

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
static int g_time_eval = 0;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

int launch_v2(const AtomArgs<double>& a_in, cudaStream_t st) {
    AtomArgs<double> a = a_in;
    {   // atoms per warp: up to kAtomsPerWarp, but never fewer than ~6 resident-block waves of blocks on the device
        // (a rank that owns 10^4 atoms keeps one atom per warp: the tail of a coarse grid costs more than the idle slots)
        int dev = 0, sms = 148;
        PANTEA_CUDA_TRY(cudaGetDevice(&dev));
        PANTEA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int64_t blocks_min = (int64_t)sms * 4 * 6;
        int apw = (int)(a.n_work / (kEvalWarps * blocks_min));
        a.apw = apw < 1 ? 1 : (apw > kAtomsPerWarp ? kAtomsPerWarp : apw);
    }
    static size_t conf_filter[64] = {0}, conf_eval[64] = {0};
    if (a.max_groups > 0) {
        const size_t smem = (size_t)kFilter2Warps * ((a.scap + 32) * sizeof(float4) + kStrip2 * sizeof(int32_t));
        int rc = opt_in_smem((const void*)pair_filter2_kernel<double>, smem, conf_filter, "pair filter: neighbour capacity too large for shared memory");
        if (rc != PANTEA_OK) return rc;
        const int blocks = (a.n_work + kFilter2Warps * kFilterAtomsPerWarp - 1) / (kFilter2Warps * kFilterAtomsPerWarp);
        pair_filter2_kernel<double><<<blocks, kFilter2Warps * 32, smem, st>>>(a);
        PANTEA_LAUNCH_CHECK();
    }
    const bool single = a.rec_bytes == kRec2fBytes;
    static size_t conf_evalf[64] = {0};
    const size_t smem = (size_t)kEvalWarps * (single ? eval2f_smem_bytes(a.scap, a.n_sf_max) : eval2_smem_bytes(a.scap, a.n_sf_max));
    int rc = single ? opt_in_smem((const void*)hdnnp_eval2f_kernel, smem, conf_evalf, "evaluation: neighbour capacity too large for shared memory")
                    : opt_in_smem((const void*)hdnnp_eval2_kernel, smem, conf_eval, "evaluation: neighbour capacity too large for shared memory");
    if (rc != PANTEA_OK) return rc;
    const int blocks = (a.n_work + kEvalWarps * a.apw - 1) / (kEvalWarps * a.apw);
    if (g_time_eval) {  // measurement hook (pantea_eval_timing): CUDA events around the dominant kernel, not capturable
        if (!g_ev0) { PANTEA_CUDA_TRY(cudaEventCreate(&g_ev0)); PANTEA_CUDA_TRY(cudaEventCreate(&g_ev1)); }
        PANTEA_CUDA_TRY(cudaEventRecord(g_ev0, st));
    }
    if (single) hdnnp_eval2f_kernel<<<blocks, kEvalWarps * 32, smem, st>>>(a);
    else hdnnp_eval2_kernel<<<blocks, kEvalWarps * 32, smem, st>>>(a);
    PANTEA_LAUNCH_CHECK();
    if (g_time_eval) PANTEA_CUDA_TRY(cudaEventRecord(g_ev1, st));
    return PANTEA_OK;
}

}  // namespace pantea

extern "C" {

// bench.py: enable != 0 brackets every later launch of the fast path's evaluation kernel with CUDA events on its stream;
// *ms (HOST, may be NULL) receives the duration of the most recent bracketed launch (synchronises on it).  Returns
// PANTEA_EINVAL when no launch has been bracketed yet.  Not usable inside stream capture.
int pantea_eval_timing(int32_t enable, float* ms) {
    using namespace pantea;
    if (ms) {
        if (!g_ev0 || !g_ev1) return fail(PANTEA_EINVAL, "pantea_eval_timing: no bracketed launch yet");
        PANTEA_CUDA_TRY(cudaEventSynchronize(g_ev1));
        PANTEA_CUDA_TRY(cudaEventElapsedTime(ms, g_ev0, g_ev1));
    }
    g_time_eval = enable ? 1 : 0;
    return PANTEA_OK;
}

}  // extern "C"
