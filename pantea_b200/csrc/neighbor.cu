// Periodic cell-list / all-pairs neighbour search for sm_100a.
//
// Replaces the reference's dense N x N masks (pantea/atoms/neighbor.py:74-115, distance.py:63-105,
// box.py:112-117) with:  counting-sort binning by cell (deterministic: atoms ordered by original index
// inside a cell), a packed, cell-ordered copy of the atom records (coalesced 16/32-byte loads), and a
// warp-cooperative candidate scan that writes neighbour rows partitioned by neighbour type.
// The neighbour predicate is evaluated with explicitly rounded, uncontracted arithmetic
// (r = sqrt((dx*dx + dy*dy) + dz*dz), 0 < r <= rc) so that the sets are bit-identical to the oracle's.
#include <cstdlib>

#include "internal.cuh"
#include "math.cuh"

namespace pantea {

constexpr int kWarpsPerBlock = 8;
#ifndef PANTEA_ROWS_MINBLOCKS
#define PANTEA_ROWS_MINBLOCKS 4
#endif
constexpr unsigned kFull = 0xffffffffu;

struct BoxArg {
    double lx, ly, lz;
    int has_box;
};

struct CellArg {
    int nx, ny, nz;
    double inv_x, inv_y, inv_z;  // cells per unit length
    int r;                       // stencil radius: 1 (cell width > list radius) or 2 (cell width > half the list radius)
};

__device__ __forceinline__ int bucket_of(int type, int n_types) { return (type >= 1 && type <= n_types) ? type - 1 : n_types; }

__device__ __forceinline__ int cell_coord(double x, double inv, int n) {
    int c = (int)floor(x * inv);
    if ((unsigned)c >= (unsigned)n) {  // outside the box: wrap (the integer division stays off the common path)
        c %= n;
        if (c < 0) c += n;
    }
    return c;
}

// ------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void cell_assign_kernel(const T* __restrict__ pos, int n, CellArg ca, BoxArg box, int32_t* __restrict__ cell_of,
                                   int32_t* __restrict__ cell_count, int32_t* __restrict__ wide_flag,
                                   const int32_t* __restrict__ guard, const uint8_t* __restrict__ role,
                                   int32_t* __restrict__ own_count, int own_begin, int own_end) {
    if (guard && *guard == 0) return;  // Verlet skin: the rows of the last rebuild are still valid
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) wide_flag[1] = 0;  // "two atoms of one cell at the same wrapped position" (set by cell_sort_pack_kernel)
    if (i >= n) return;
    if (role && role[i] == 0) { cell_of[i] = -1; return; }  // not on this rank (brick decomposition)
    {   // the screening pass measures true minimum-image distances; they equal the reference's single-shift ones only
        // while every coordinate difference stays below 1.5 box lengths
        const double x = (double)pos[3 * i], y = (double)pos[3 * i + 1], z = (double)pos[3 * i + 2];
        const bool wide = !(x >= -0.25 * box.lx && x <= 1.25 * box.lx && y >= -0.25 * box.ly && y <= 1.25 * box.ly &&
                            z >= -0.25 * box.lz && z <= 1.25 * box.lz);
        if (wide) *wide_flag = 1;
    }
    int cx = cell_coord((double)pos[3 * i], ca.inv_x, ca.nx);
    int cy = cell_coord((double)pos[3 * i + 1], ca.inv_y, ca.ny);
    int cz = cell_coord((double)pos[3 * i + 2], ca.inv_z, ca.nz);
    int c = (cz * ca.ny + cy) * ca.nx + cx;
    cell_of[i] = c;
    atomicAdd(&cell_count[c], 1);
    // owned atoms per cell (ranks that own a part of the atoms): scanned together with the cell counts
    if (own_count && (role ? role[i] == 2 : (i >= own_begin && i < own_end))) atomicAdd(&own_count[c], 1);
}

// Exclusive scan of counts[0..n) into start[0..n] (start[n] = total); also zeroes `fill`.  Two launches: every block
// of 256 threads sums its tile of 2048 cells, then every block scans its own tile behind the sum of the tiles before
// it (a single block took 38 us for the 3 x 10^4 half-width cells of the 10^5-atom box).  counts may alias start or
// fill: a thread reads all of its inputs before it writes its outputs, and the tile sums are complete before any write.
constexpr int kScanThreads = 256, kScanK = 8, kScanTile = kScanThreads * kScanK;

__global__ void __launch_bounds__(kScanThreads) cell_tile_sums_kernel(const int32_t* __restrict__ counts, int n,
                                                                      int32_t* __restrict__ tile_sums,
                                                                      const int32_t* __restrict__ guard,
                                                                      const int32_t* __restrict__ counts_b) {
    if (guard && *guard == 0) return;
    if (blockIdx.y == 1) { counts = counts_b; tile_sums += gridDim.x; }  // second scan of the same launch
    __shared__ int warp_sums[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int i0 = blockIdx.x * kScanTile + tid * kScanK;
    int x = 0;
#pragma unroll
    for (int k = 0; k < kScanK; ++k) x += i0 + k < n ? counts[i0 + k] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if (lane == 0) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
        if (lane == 0) tile_sums[blockIdx.x] = w;
    }
}

__global__ void __launch_bounds__(kScanThreads) cell_scan_kernel(const int32_t* counts, int n, const int32_t* __restrict__ tile_sums,
                                                                 int32_t* start, int32_t* fill, const int32_t* __restrict__ guard,
                                                                 const int32_t* counts_b, int32_t* start_b, int32_t* fill_b) {
    if (guard && *guard == 0) return;
    if (blockIdx.y == 1) { counts = counts_b; start = start_b; fill = fill_b; tile_sums += gridDim.x; }
    __shared__ int warp_sums[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // sum of the tiles before this one (a few dozen values at most: one warp)
    if (wid == 0) {
        int b = 0;
        for (int t = lane; t < (int)blockIdx.x; t += 32) b += tile_sums[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(kFull, b, o);
        if (lane == 0) base_s = b;
    }
    const int i0 = blockIdx.x * kScanTile + tid * kScanK;
    int v[kScanK];
    int x = 0;
#pragma unroll
    for (int k = 0; k < kScanK; ++k) {
        v[k] = i0 + k < n ? counts[i0 + k] : 0;
        x += v[k];
    }
    const int mine = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, w, o);
            if (lane >= o) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const int incl = x + (wid > 0 ? warp_sums[wid - 1] : 0) + base_s;
    int run = incl - mine;
#pragma unroll
    for (int k = 0; k < kScanK; ++k) {
        if (i0 + k < n) {
            start[i0 + k] = run;
            fill[i0 + k] = 0;
        }
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && tid == kScanThreads - 1) start[n] = incl;
}

// scans counts -> start (zeroing fill) and, when counts_b is given, counts_b -> start_b (zeroing fill_b) in the same launches
static int launch_cell_scan(pantea_workspace* ws, const int32_t* counts, int n, int32_t* start, int32_t* fill,
                            const int32_t* counts_b, int32_t* start_b, int32_t* fill_b, const int32_t* guard, cudaStream_t st) {
    const int tiles = (n + kScanTile - 1) / kScanTile;
    const dim3 grid(tiles, counts_b ? 2 : 1);
    cell_tile_sums_kernel<<<grid, kScanThreads, 0, st>>>(counts, n, ws->scan_sums, guard, counts_b);
    PANTEA_LAUNCH_CHECK();
    cell_scan_kernel<<<grid, kScanThreads, 0, st>>>(counts, n, ws->scan_sums, start, fill, guard, counts_b, start_b, fill_b);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

__global__ void cell_scatter_kernel(const int32_t* __restrict__ cell_of, int n, const int32_t* __restrict__ cell_start,
                                    int32_t* __restrict__ cell_fill, int32_t* __restrict__ tmp_order,
                                    const int32_t* __restrict__ guard) {
    if (guard && *guard == 0) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    if (c < 0) return;
    int p = cell_start[c] + atomicAdd(&cell_fill[c], 1);
    tmp_order[p] = i;
}

// one warp per cell: order the cell's atoms by original index (rank sort) and emit the packed records
template <typename T>
__global__ void cell_sort_pack_kernel(const T* __restrict__ pos, const int32_t* __restrict__ types, int n_types,
                                      const int32_t* __restrict__ cell_start, int ncells,
                                      const int32_t* __restrict__ tmp_order, Rec<T>* __restrict__ rec,
                                      int32_t* __restrict__ slot_of, Rec<float>* __restrict__ rec_screen, BoxArg box,
                                      int32_t* __restrict__ cell_fill, const int32_t* __restrict__ guard,
                                      int own_begin, int own_end, int32_t* __restrict__ cell_own,
                                      int32_t* __restrict__ tcount, const uint8_t* __restrict__ role,
                                      int32_t* __restrict__ dup_flag) {
    if (guard && *guard == 0) return;
    auto is_owned = [&](int idx) { return role ? role[idx] == 2 : (idx >= own_begin && idx < own_end); };
    int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (cell >= ncells) return;
    if (lane == 0) cell_fill[cell] = 0;  // leave the counting-sort scratch zeroed for the next (device-decided) rebuild
    int lo = cell_start[cell], hi = cell_start[cell + 1];
    for (int a = lo + lane; a < hi; a += 32) {
        int mine = tmp_order[a];
        int rank = 0;
        for (int b = lo; b < hi; ++b) rank += (tmp_order[b] < mine) ? 1 : 0;
        Rec<T> r;
        r.x = pos[3 * mine]; r.y = pos[3 * mine + 1]; r.z = pos[3 * mine + 2];
        rec_set(r, bucket_of(types[mine], n_types), mine);
        rec[lo + rank] = r;
        slot_of[mine] = lo + rank;
        if (cell_own && !is_owned(mine))  // no row is built for it: report zero neighbours
            for (int b = 0; b < kBuckets; ++b) tcount[(size_t)(lo + rank) * kBuckets + b] = 0;
        if (rec_screen) {  // box-wrapped coordinates in [0, L], rounded to float
            const double x = (double)r.x, y = (double)r.y, z = (double)r.z;
            Rec<float> f;
            f.x = (float)(x - box.lx * floor(x / box.lx));
            f.y = (float)(y - box.ly * floor(y / box.ly));
            f.z = (float)(z - box.lz * floor(z / box.lz));
            rec_set(f, bucket_of(types[mine], n_types), mine);
            rec_screen[lo + rank] = f;
            // Coincident atoms (the reference drops a triplet whose two neighbours sit at the same position, acsf.py:325)
            // share a cell and their wrapped single-precision coordinates: raise the flag that sends the pair filter
            // through its exact test.  Conservative: equal floats / more than 32 atoms in a cell are enough.
            if (hi - lo > 32) *dup_flag = 1;
            const unsigned act = __activemask();
            const int cnt = min(32, hi - (a - lane));
            for (int b = 0; b < cnt; ++b) {
                const float ox = __shfl_sync(act, f.x, b), oy = __shfl_sync(act, f.y, b), oz = __shfl_sync(act, f.z, b);
                if (b != lane && ox == f.x && oy == f.y && oz == f.z) *dup_flag = 1;
            }
        }
    }
}

// one warp per cell: slots of the cell's owned atoms, in slot order, at own_start[cell] of the compact list
template <typename T>
__global__ void owned_fill_kernel(const Rec<T>* __restrict__ rec, const int32_t* __restrict__ cell_start, int ncells,
                                  const int32_t* __restrict__ own_start, int own_begin, int own_end,
                                  int32_t* __restrict__ owned_slots, const int32_t* __restrict__ guard,
                                  const uint8_t* __restrict__ role) {
    if (guard && *guard == 0) return;
    int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (cell >= ncells) return;
    const int lo = cell_start[cell], hi = cell_start[cell + 1];
    int out = own_start[cell];
    for (int s0 = lo; s0 < hi; s0 += 32) {
        const int s = s0 + lane;
        bool mine = false;
        if (s < hi) {
            const int idx = rec_idx(rec[s]);
            mine = role ? role[idx] == 2 : (idx >= own_begin && idx < own_end);
        }
        const unsigned m = __ballot_sync(kFull, mine);
        if (mine) owned_slots[out + __popc(m & ((1u << lane) - 1u))] = s;
        out += __popc(m);
    }
}

// all-pairs mode: records keep the original order
template <typename T>
__global__ void pack_identity_kernel(const T* __restrict__ pos, const int32_t* __restrict__ types, int n_types, int n,
                                     const int32_t* __restrict__ struct_ptr, int n_structs, Rec<T>* __restrict__ rec,
                                     int32_t* __restrict__ slot_of, int32_t* __restrict__ struct_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rec<T> r;
    r.x = pos[3 * i]; r.y = pos[3 * i + 1]; r.z = pos[3 * i + 2];
    rec_set(r, bucket_of(types[i], n_types), i);
    rec[i] = r;
    slot_of[i] = i;
    int s = 0;
    if (struct_ptr) {  // last s with struct_ptr[s] <= i
        int lo = 0, hi = n_structs;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (struct_ptr[mid] <= i) lo = mid; else hi = mid;
        }
        s = lo;
    }
    struct_of[i] = s;
}

// ------------------------------------------------------------------------------------------------
// neighbour rows: one warp per atom (cell-sorted slot)
// ------------------------------------------------------------------------------------------------
struct RowArgs {
    int n, cap, n_buckets;
    BoxArg box;
    CellArg cell;
    const int32_t* cell_start;
    const int32_t* struct_of;
    const int32_t* struct_ptr;
    const double* boxes;
    double rc;
    int own_begin, own_end;  // original-index ownership range
    int32_t* nbr;
    int32_t* tcount;
    int32_t* flags;
    const Rec<float>* rec_screen;  // FP32 screening records (F64 cell mode) or NULL
    const int32_t* wide_flag;
    float screen_band;             // |r2_f32 - r2| bound around rc^2 (and around 0)
    const int32_t* guard;          // Verlet skin: skip the kernel while *guard == 0 (NULL: always run)
    const int32_t* owned_slots;    // block-owned ranks: the warps walk this compact slot list (NULL: every slot)
    int n_owned;
    int tmp_cap;                   // per-warp scratch entries behind the row lists (0: none)
    const int32_t* n_owned_dev;    // role mode: the owned count lives on the device (NULL: n_owned)
    const uint8_t* role;           // role mode: every slot of the owned list is owned
};

template <typename T, int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PANTEA_ROWS_MINBLOCKS) neighbor_rows_kernel(const Rec<T>* __restrict__ rec, RowArgs a) {
    extern __shared__ int32_t smem_rows[];
    if (a.guard && *a.guard == 0) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= (a.owned_slots ? (a.n_owned_dev ? *a.n_owned_dev : a.n_owned) : a.n)) return;
    const int i = a.owned_slots ? a.owned_slots[w] : w;
    int32_t* L = smem_rows + wib * a.cap;
    const Rec<T> ri = rec[i];
    const int oi = rec_idx(ri);
    if (!a.role && (oi < a.own_begin || oi >= a.own_end)) {
        if (lane < kBuckets) a.tcount[(size_t)i * kBuckets + lane] = 0;
        return;
    }
    T lx = (T)a.box.lx, ly = (T)a.box.ly, lz = (T)a.box.lz;
    bool pbc = a.box.has_box != 0;
    int lo_all = 0, hi_all = a.n;
    if (MODE == kModeAllPairs) {
        int s = a.struct_of[i];
        if (a.struct_ptr) { lo_all = a.struct_ptr[s]; hi_all = a.struct_ptr[s + 1]; }
        if (a.boxes) { lx = (T)a.boxes[3 * s]; ly = (T)a.boxes[3 * s + 1]; lz = (T)a.boxes[3 * s + 2]; pbc = true; }
    }
    const T rc = (T)a.rc;
    const T rel = sizeof(T) == 8 ? (T)1e-12 : (T)1e-5;
    const T rc2_lo = rc * rc * ((T)1 - rel), rc2_hi = rc * rc * ((T)1 + rel);
    int cnt = 0;

    auto scan = [&](int lo, int hi) {
        for (int j0 = lo; j0 < hi; j0 += 32) {
            int j = j0 + lane;
            bool ok = false;
            int bucket = 0;
            if (j < hi) {
                Rec<T> rj = rec[j];
                T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
                if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
                // predicate (r <= rc) & (r > 0) on r = sqrt((dx^2 + dy^2) + dz^2): the correctly rounded root is only
                // needed when r^2 is within rounding distance of rc^2
                const T r2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
                if (r2 < rc2_lo) ok = r2 > (T)0;
                else if (r2 > rc2_hi) ok = false;
                else { const T r = norm3_rn(dx, dy, dz); ok = (r <= rc) && (r > (T)0); }
                bucket = rec_type(rj);
            }
            unsigned m = __ballot_sync(kFull, ok);
            int p = cnt + __popc(m & ((1u << lane) - 1u));
            if (ok && p < a.cap) L[p] = j | (bucket << 28);
            cnt += __popc(m);
        }
    };

    // FP32 screening (F64 workspaces, cell mode): |d| per axis is min(|dx|, L - |dx|) on box-wrapped single-precision
    // coordinates; only candidates whose r^2 falls within the error band of rc^2 (or of 0) take the exact path above.
    // The accepted set is therefore identical to the exact predicate's.
    const bool screen = sizeof(T) == 8 && MODE == kModeCell && a.rec_screen != nullptr && *a.wide_flag == 0;
    Rec<float> rif;
    rif.x = rif.y = rif.z = 0.f;
    if (screen) rif = a.rec_screen[i];
    const float flx = (float)a.box.lx, fly = (float)a.box.ly, flz = (float)a.box.lz;
    const float rc2f = (float)(a.rc * a.rc), lo_f = rc2f - a.screen_band, hi_f = rc2f + a.screen_band;
    auto scan_screen = [&](int lo, int hi) {
        for (int j0 = lo; j0 < hi; j0 += 32) {
            const int j = j0 + lane;
            bool ok = false;
            int bucket = 0;
            if (j < hi) {
                const Rec<float> rj = a.rec_screen[j];
                float ax = fabsf(rif.x - rj.x), ay = fabsf(rif.y - rj.y), az = fabsf(rif.z - rj.z);
                ax = fminf(ax, flx - ax); ay = fminf(ay, fly - ay); az = fminf(az, flz - az);
                const float r2f = ax * ax + ay * ay + az * az;
                ok = r2f < lo_f;
                const bool ambiguous = (ok ? r2f <= a.screen_band : r2f <= hi_f) && j != i;
                if (j == i) ok = false;
                if (ambiguous) {
                    const Rec<T> rj = rec[j];
                    T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
                    dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz);
                    const T r = norm3_rn(dx, dy, dz);
                    ok = (r <= rc) && (r > (T)0);
                }
                bucket = rec_type(rj);
            }
            const unsigned m = __ballot_sync(kFull, ok);
            const int p = cnt + __popc(m & ((1u << lane) - 1u));
            if (ok && p < a.cap) L[p] = j | (bucket << 28);
            cnt += __popc(m);
        }
    };
    auto scan_any = [&](int lo, int hi) {
        if (screen) scan_screen(lo, hi); else scan(lo, hi);
    };

    if (MODE == kModeCell && a.cell.r == 2 && screen && a.tmp_cap > 0) {
        // Half-width cells, 5 x 5 x 5 stencil.  Lane r < 25 owns the (dy, dz) stencil row r and writes the slots of its
        // candidates into its stretch of a per-warp scratch list; the flattened list -- in exactly the order of the scan
        // below (dz outer, dy inner, wrapped x-part first, slots ascending) -- is then walked by all 32 lanes with one
        // shared-memory read per candidate instead of a five-step shuffle search.  The x-range of a row is
        // clipped to the cells that can hold an atom within the list radius given the row's (dy, dz) offset and the
        // atom's position inside its own cell: a dropped cell lies farther than rc in the periodic metric, hence also
        // in the reference's single-shift metric, so the accepted set does not change.
        const int nx = a.cell.nx, ny = a.cell.ny, nz = a.cell.nz;
        const double ux = (double)ri.x * a.cell.inv_x, uy = (double)ri.y * a.cell.inv_y, uz = (double)ri.z * a.cell.inv_z;
        const double fx0 = floor(ux), fy0 = floor(uy), fz0 = floor(uz);
        const int cx = cell_coord((double)ri.x, a.cell.inv_x, nx);
        const int cy = cell_coord((double)ri.y, a.cell.inv_y, ny);
        const int cz = cell_coord((double)ri.z, a.cell.inv_z, nz);
        int a0 = 0, la = 0, b0 = 0, lb = 0;
        if (lane < 25) {
            const int oz = lane / 5 - 2, oy = lane % 5 - 2;
            int z = cz + oz, y = cy + oy;
            z = z < 0 ? z + nz : (z >= nz ? z - nz : z);
            y = y < 0 ? y + ny : (y >= ny ? y - ny : y);
            // smallest possible |dy|, |dz| to any point of the row's cells (in length units)
            const double gy = oy == 0 ? 0.0 : (oy > 0 ? (double)oy - (uy - fy0) : (uy - fy0) - (double)(oy + 1)) / a.cell.inv_y;
            const double gz = oz == 0 ? 0.0 : (oz > 0 ? (double)oz - (uz - fz0) : (uz - fz0) - (double)(oz + 1)) / a.cell.inv_z;
            const double budget = a.rc * a.rc * (1.0 + 1e-9) - gy * gy - gz * gz;
            if (budget >= 0.0) {
                const double hx = (sqrt(budget) + 1e-9 * a.rc) * a.cell.inv_x;  // reach along x in cell units
                int xlo = (int)floor(ux - hx) - (int)fx0, xhi = (int)floor(ux + hx) - (int)fx0;  // relative to the own cell
                xlo = xlo < -2 ? -2 : xlo; xhi = xhi > 2 ? 2 : xhi;
                xlo += cx; xhi += cx;
                const int rowc = (z * ny + y) * nx;
                if (xlo >= 0 && xhi < nx) {
                    a0 = a.cell_start[rowc + xlo]; la = a.cell_start[rowc + xhi + 1] - a0;
                } else if (xlo < 0) {  // cells xlo + nx .. nx - 1, then 0 .. xhi
                    a0 = a.cell_start[rowc + xlo + nx]; la = a.cell_start[rowc + nx] - a0;
                    b0 = a.cell_start[rowc]; lb = a.cell_start[rowc + xhi + 1] - b0;
                } else {               // cells xlo .. nx - 1, then 0 .. xhi - nx
                    a0 = a.cell_start[rowc + xlo]; la = a.cell_start[rowc + nx] - a0;
                    b0 = a.cell_start[rowc]; lb = a.cell_start[rowc + xhi - nx + 1] - b0;
                }
            }
        }
        const int mine = la + lb;
        int pe = mine;  // inclusive scan over the lanes: every lane's stretch of the scratch list
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, pe, o);
            if (lane >= o) pe += t;
        }
        const int n_cand = __shfl_sync(kFull, pe, 31);
        int32_t* tmp = smem_rows + kWarpsPerBlock * a.cap + wib * a.tmp_cap;
        if (n_cand <= a.tmp_cap) {
            // every lane writes out the slots of its row (a dozen shared-memory stores), then all 32 lanes walk the
            // flattened list: one LDS per candidate instead of a shuffle search, two 32-wide steps of loads in flight
            int32_t* mt = tmp + (pe - mine);
            for (int k = 0; k < mine; ++k) mt[k] = k < la ? a0 + k : b0 + (k - la);
            __syncwarp();
#ifndef PANTEA_ROWS_RU
#define PANTEA_ROWS_RU 2
#endif
            constexpr int RU = PANTEA_ROWS_RU;
            for (int v0 = 0; v0 < n_cand; v0 += 32 * RU) {
                int jj[RU];
                Rec<float> rq[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int v = v0 + 32 * u + lane;
                    jj[u] = v < n_cand ? tmp[v] : i;  // past the end: the atom itself (rejected)
                    rq[u] = a.rec_screen[jj[u]];
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int j = jj[u];
                    const Rec<float> rj = rq[u];
                    float ax = fabsf(rif.x - rj.x), ay = fabsf(rif.y - rj.y), az = fabsf(rif.z - rj.z);
                    ax = fminf(ax, flx - ax); ay = fminf(ay, fly - ay); az = fminf(az, flz - az);
                    const float r2f = ax * ax + ay * ay + az * az;
                    bool ok = r2f < lo_f && j != i;
                    const bool ambiguous = (ok ? r2f <= a.screen_band : r2f <= hi_f) && j != i;
                    if (ambiguous) {
                        const Rec<T> rjx = rec[j];
                        T dx = sub_rn(ri.x, rjx.x), dy = sub_rn(ri.y, rjx.y), dz = sub_rn(ri.z, rjx.z);
                        dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz);
                        const T rr = norm3_rn(dx, dy, dz);
                        ok = (rr <= rc) && (rr > (T)0);
                    }
                    const unsigned m = __ballot_sync(kFull, ok);
                    const int p = cnt + __popc(m & ((1u << lane) - 1u));
                    if (ok && p < a.cap) L[p] = j | (rec_type(rj) << 28);
                    cnt += __popc(m);
                }
            }
        } else {
            cnt = a.cap + 1;  // scratch too small for this many candidates: report as an overflow (the caller re-runs)
        }
    } else if (MODE == kModeCell && a.cell.r == 2) {
        // Half-width cells, 5 x 5 x 5 stencil: 125 (w/2)^3 = 15.6 w^3 of candidate volume instead of 27 w^3.  A (dy, dz)
        // row of the stencil now holds only ~17 candidates, so the 25 rows are not scanned one by one (half-empty
        // warps) but as ONE flattened candidate list: lane r < 25 owns row r -- up to two contiguous slot ranges, two
        // when the x-stencil wraps around the box -- an inclusive warp scan of the range lengths gives every row its
        // offset, and each 32-wide step finds the row of its candidate by a 5-step search over the lanes' offsets.
        // Row order (dz outer, dy inner, wrapped x-part first) is the order of the 3 x 3 x 3 scan below.
        const int nx = a.cell.nx, ny = a.cell.ny, nz = a.cell.nz;
        const int cx = cell_coord((double)ri.x, a.cell.inv_x, nx);
        const int cy = cell_coord((double)ri.y, a.cell.inv_y, ny);
        const int cz = cell_coord((double)ri.z, a.cell.inv_z, nz);
        int a0 = 0, la = 0, b0 = 0, lb = 0;
        if (lane < 25) {
            int z = cz + lane / 5 - 2, y = cy + lane % 5 - 2;
            z = z < 0 ? z + nz : (z >= nz ? z - nz : z);
            y = y < 0 ? y + ny : (y >= ny ? y - ny : y);
            const int rowc = (z * ny + y) * nx;
            const int xlo = cx - 2, xhi = cx + 2;
            if (xlo >= 0 && xhi < nx) {
                a0 = a.cell_start[rowc + xlo]; la = a.cell_start[rowc + xhi + 1] - a0;
            } else if (xlo < 0) {  // cells xlo + nx .. nx - 1, then 0 .. xhi
                a0 = a.cell_start[rowc + xlo + nx]; la = a.cell_start[rowc + nx] - a0;
                b0 = a.cell_start[rowc]; lb = a.cell_start[rowc + xhi + 1] - b0;
            } else {               // cells xlo .. nx - 1, then 0 .. xhi - nx
                a0 = a.cell_start[rowc + xlo]; la = a.cell_start[rowc + nx] - a0;
                b0 = a.cell_start[rowc]; lb = a.cell_start[rowc + xhi - nx + 1] - b0;
            }
        }
        int pe = la + lb;  // inclusive scan over the lanes: end offset of every row in the flattened list
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, pe, o);
            if (lane >= o) pe += t;
        }
        const int ps = pe - (la + lb);
        const int n_cand = __shfl_sync(kFull, pe, 31);
        for (int v0 = 0; v0 < n_cand; v0 += 32) {
            const int v = v0 + lane;
            int r = 0;  // number of rows that end at or before v == the row of candidate v
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const int pv = __shfl_sync(kFull, pe, (r + st - 1) & 31);
                if (pv <= v) r += st;
            }
            r &= 31;
            const int off = v - __shfl_sync(kFull, ps, r);
            const int ra0 = __shfl_sync(kFull, a0, r), rla = __shfl_sync(kFull, la, r), rb0 = __shfl_sync(kFull, b0, r);
            const int j = off < rla ? ra0 + off : rb0 + (off - rla);
            const bool has = v < n_cand;
            bool ok = false;
            int bucket = 0;
            if (has) {
                if (screen) {
                    const Rec<float> rj = a.rec_screen[j];
                    float ax = fabsf(rif.x - rj.x), ay = fabsf(rif.y - rj.y), az = fabsf(rif.z - rj.z);
                    ax = fminf(ax, flx - ax); ay = fminf(ay, fly - ay); az = fminf(az, flz - az);
                    const float r2f = ax * ax + ay * ay + az * az;
                    ok = r2f < lo_f;
                    const bool ambiguous = (ok ? r2f <= a.screen_band : r2f <= hi_f) && j != i;
                    if (j == i) ok = false;
                    if (ambiguous) {
                        const Rec<T> rjx = rec[j];
                        T dx = sub_rn(ri.x, rjx.x), dy = sub_rn(ri.y, rjx.y), dz = sub_rn(ri.z, rjx.z);
                        dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz);
                        const T rr = norm3_rn(dx, dy, dz);
                        ok = (rr <= rc) && (rr > (T)0);
                    }
                    bucket = rec_type(rj);
                } else {
                    const Rec<T> rj = rec[j];
                    T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
                    if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
                    const T r2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
                    if (r2 < rc2_lo) ok = r2 > (T)0;
                    else if (r2 > rc2_hi) ok = false;
                    else { const T rr = norm3_rn(dx, dy, dz); ok = (rr <= rc) && (rr > (T)0); }
                    bucket = rec_type(rj);
                }
            }
            const unsigned m = __ballot_sync(kFull, ok);
            const int p = cnt + __popc(m & ((1u << lane) - 1u));
            if (ok && p < a.cap) L[p] = j | (bucket << 28);
            cnt += __popc(m);
        }
    } else if (MODE == kModeCell) {
        int cx = cell_coord((double)ri.x, a.cell.inv_x, a.cell.nx);
        int cy = cell_coord((double)ri.y, a.cell.inv_y, a.cell.ny);
        int cz = cell_coord((double)ri.z, a.cell.inv_z, a.cell.nz);
        for (int dz = -1; dz <= 1; ++dz) {
            int z = cz + dz; z = z < 0 ? z + a.cell.nz : (z >= a.cell.nz ? z - a.cell.nz : z);
            for (int dy = -1; dy <= 1; ++dy) {
                int y = cy + dy; y = y < 0 ? y + a.cell.ny : (y >= a.cell.ny ? y - a.cell.ny : y);
                int rowc = (z * a.cell.ny + y) * a.cell.nx;
                // the three x-cells are contiguous in memory unless the stencil wraps around
                if (cx > 0 && cx < a.cell.nx - 1) {
                    scan_any(a.cell_start[rowc + cx - 1], a.cell_start[rowc + cx + 2]);
                } else {
                    for (int dx = -1; dx <= 1; ++dx) {
                        int x = cx + dx; x = x < 0 ? x + a.cell.nx : (x >= a.cell.nx ? x - a.cell.nx : x);
                        scan_any(a.cell_start[rowc + x], a.cell_start[rowc + x + 1]);
                    }
                }
            }
        }
    } else {
        scan(lo_all, hi_all);
    }
    __syncwarp();
    if (lane == 0) atomicMax(&a.flags[0], cnt);
    const int total = cnt < a.cap ? cnt : a.cap;

    // bucket counts, then stable placement partitioned by bucket
    int c[kBuckets];
#pragma unroll
    for (int b = 0; b < kBuckets; ++b) c[b] = 0;
    for (int e0 = 0; e0 < total; e0 += 32) {
        int e = e0 + lane;
        int bk = e < total ? (int)((unsigned)L[e] >> 28) : -1;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) {
            if (b >= a.n_buckets) break;  // warp-uniform: only the potential's element buckets (+ "other") are in use
            c[b] += __popc(__ballot_sync(kFull, bk == b));
        }
    }
    int run[kBuckets];
    int acc = 0;
#pragma unroll
    for (int b = 0; b < kBuckets; ++b) { run[b] = acc; acc += c[b]; }
    int32_t* row = a.nbr + (size_t)i * a.cap;
    for (int e0 = 0; e0 < total; e0 += 32) {
        int e = e0 + lane;
        int v = e < total ? L[e] : 0;
        int bk = e < total ? (int)((unsigned)v >> 28) : -1;
        int dst = -1;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) {
            if (b >= a.n_buckets) break;
            unsigned m = __ballot_sync(kFull, bk == b);
            if (bk == b) dst = run[b] + __popc(m & ((1u << lane) - 1u));
            run[b] += __popc(m);
        }
        if (dst >= 0) row[dst] = v & 0x0fffffff;
    }
#pragma unroll
    for (int b = 0; b < kBuckets; ++b)
        if (lane == b) a.tcount[(size_t)i * kBuckets + b] = c[b];
}

// ------------------------------------------------------------------------------------------------
// Verlet skin: rebuild decision, record refresh, bookkeeping -- all on the device, so that the launch sequence of a
// pantea_neighbor_build call does not depend on the outcome (CUDA-graph safe)
// ------------------------------------------------------------------------------------------------
// flags[0] = 1 when some atom moved more than skin / 2 (minimum image: the integrator wraps positions) since the rebuild
template <typename T>
__global__ void skin_check_kernel(const T* __restrict__ pos, const T* __restrict__ ref, int n, BoxArg box, double limit2,
                                  int32_t* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dx = (double)pos[3 * i] - (double)ref[3 * i], dy = (double)pos[3 * i + 1] - (double)ref[3 * i + 1],
           dz = (double)pos[3 * i + 2] - (double)ref[3 * i + 2];
    dx -= box.lx * rint(dx / box.lx); dy -= box.ly * rint(dy / box.ly); dz -= box.lz * rint(dz / box.lz);
    if (!(dx * dx + dy * dy + dz * dz <= limit2)) flags[0] = 1;  // also catches NaN
}

// the packed records follow the atoms every call (the slot order of the last rebuild stays)
template <typename T>
__global__ void skin_refresh_kernel(const T* __restrict__ pos, Rec<T>* __restrict__ rec, int n) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    Rec<T> r = rec[s];
    const int i = rec_idx(r);
    r.x = pos[3 * i]; r.y = pos[3 * i + 1]; r.z = pos[3 * i + 2];
    rec[s] = r;
}

template <typename T>
__global__ void skin_save_kernel(const T* __restrict__ pos, T* __restrict__ ref, int n3, const int32_t* __restrict__ guard) {
    if (*guard == 0) return;
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n3) ref[e] = pos[e];
}

__global__ void skin_finish_kernel(int32_t* __restrict__ flags, int32_t* __restrict__ wide_flag) {
    if (flags[0] != 0) {
        flags[1] = 1;  // the pair lists of the evaluation are stale
        flags[2] += 1;
        *wide_flag = 0;  // consumed by the rows kernel of this rebuild
    }
    flags[0] = 0;
    flags[3] += 1;
}

// ------------------------------------------------------------------------------------------------
// export helpers
// ------------------------------------------------------------------------------------------------
__global__ void neighbor_counts_kernel(const int32_t* __restrict__ slot_of, const int32_t* __restrict__ tcount, int n,
                                       int32_t* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* t = tcount + (size_t)slot_of[i] * kBuckets;
    int s = 0;
#pragma unroll
    for (int b = 0; b < kBuckets; ++b) s += t[b];
    counts[i] = s;
}

template <typename T>
__global__ void neighbor_export_kernel(const Rec<T>* __restrict__ rec, const int32_t* __restrict__ slot_of,
                                       const int32_t* __restrict__ tcount, const int32_t* __restrict__ nbr, int cap, int n,
                                       const int64_t* __restrict__ row_ptr, int32_t* __restrict__ col) {
    extern __shared__ int32_t smem_exp[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * kWarpsPerBlock + wib;
    if (i >= n) return;
    int32_t* ids = smem_exp + wib * cap;
    const int slot = slot_of[i];
    int cnt = 0;
#pragma unroll
    for (int b = 0; b < kBuckets; ++b) cnt += tcount[(size_t)slot * kBuckets + b];
    if (cnt > cap) cnt = cap;
    for (int e = lane; e < cnt; e += 32) ids[e] = rec_idx(rec[nbr[(size_t)slot * cap + e]]);
    __syncwarp();
    const int64_t base = row_ptr[i];
    for (int e = lane; e < cnt; e += 32) {
        int mine = ids[e], rank = 0;
        for (int q = 0; q < cnt; ++q) rank += ids[q] < mine ? 1 : 0;
        col[base + rank] = mine;
    }
}

template <typename T>
__global__ void distances_kernel(const Rec<T>* __restrict__ rec, const int32_t* __restrict__ slot_of,
                                 const int32_t* __restrict__ struct_of, const double* __restrict__ boxes, BoxArg box,
                                 const int32_t* __restrict__ idx_i, int64_t n_i, const int32_t* __restrict__ idx_j,
                                 int64_t n_j, T* __restrict__ r_out, T* __restrict__ d_out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_i * n_j) return;
    int64_t a = t / n_j, b = t - a * n_j;
    int si = slot_of[idx_i ? idx_i[a] : (int)a], sj = slot_of[idx_j ? idx_j[b] : (int)b];
    Rec<T> ri = rec[si], rj = rec[sj];
    T lx = (T)box.lx, ly = (T)box.ly, lz = (T)box.lz;
    bool pbc = box.has_box != 0;
    if (boxes) { int s = struct_of[si]; lx = (T)boxes[3 * s]; ly = (T)boxes[3 * s + 1]; lz = (T)boxes[3 * s + 2]; pbc = true; }
    T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
    if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
    r_out[t] = norm3_rn(dx, dy, dz);
    if (d_out) { d_out[3 * t] = dx; d_out[3 * t + 1] = dy; d_out[3 * t + 2] = dz; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename T>
static int build_typed(pantea_workspace* ws, const void* pos_v, const int32_t* types, int64_t n, const double* box,
                       const int32_t* struct_ptr, const double* boxes, int64_t n_structs, double rc, cudaStream_t st) {
    const T* pos = (const T*)pos_v;
    Rec<T>* rec = (Rec<T>*)ws->rec;
    const int threads = 256;
    const int blocks_n = (int)((n + threads - 1) / threads);
    ws->n = n;
    ws->rc = rc;
    ws->has_box = box != nullptr;
    ws->struct_ptr = struct_ptr;
    ws->boxes = boxes;
    ws->n_structs = struct_ptr ? n_structs : 1;
    for (int k = 0; k < 3; ++k) ws->box[k] = box ? box[k] : 0.0;
    BoxArg ba{ws->box[0], ws->box[1], ws->box[2], ws->has_box ? 1 : 0};
    CellArg ca{1, 1, 1, 0, 0, 0, 1};

    bool use_cells = false, use_skin = false;
    if (box && !struct_ptr) {
        // cell width strictly larger than the list radius (relative margin covers the rounding of x * inv); a Verlet
        // skin widens the radius when the box still holds three cells per axis, otherwise it is ignored
        int nc[3];
        double w_list = 0.0;
        for (int pass = ws->skin > 0.0 ? 0 : 1; pass < 2 && !use_cells; ++pass) {
            const double w = (pass == 0 ? rc + ws->skin : rc) * (1.0 + 1e-7);
            for (int k = 0; k < 3; ++k) nc[k] = (int)std::floor(box[k] / w);
            use_cells = nc[0] >= 3 && nc[1] >= 3 && nc[2] >= 3;
            use_skin = use_cells && pass == 0;
            w_list = w;
        }
        if (use_skin) rc += ws->skin;
        ws->rc = rc;
        if (use_cells) {
            // half-width cells with a 5 x 5 x 5 stencil when the box holds five of them per axis: 42 % less candidate
            // volume (two atoms closer than the list radius are at most two such cells apart along every axis).
            // PANTEA_CELL_STENCIL=1 in the environment keeps the 3 x 3 x 3 stencil (diagnostics).
            static const bool allow_fine = []() { const char* e = std::getenv("PANTEA_CELL_STENCIL"); return !(e && e[0] == '1'); }();
            int stencil = 1;
            // the choice follows the SYSTEM size (not a rank's share): the row order, hence every bit of the result,
            // is then the same on 1 and on N GPUs; small systems keep the coarse grid (fewer cells to scan and sort)
            if (allow_fine && n >= 2048) {
                int nf[3];
                for (int k = 0; k < 3; ++k) nf[k] = (int)std::floor(box[k] / (0.5 * w_list));
                if (nf[0] >= 5 && nf[1] >= 5 && nf[2] >= 5) { stencil = 2; for (int k = 0; k < 3; ++k) nc[k] = nf[k]; }
            }
            const int nc_min = stencil == 2 ? 5 : 3;
            // keep the cell count bounded for sparse systems / tiny cutoffs: at most ~4 cells per atom (fewer, wider
            // cells keep the stencil sufficient)
            for (int k = 0; k < 3; ++k) if (nc[k] > 1024) nc[k] = 1024;
            const int64_t max_cells = 4 * n > 64 ? 4 * n : 64;
            while ((int64_t)nc[0] * nc[1] * nc[2] > max_cells) {
                int big = nc[0] >= nc[1] ? (nc[0] >= nc[2] ? 0 : 2) : (nc[1] >= nc[2] ? 1 : 2);
                if (nc[big] <= nc_min) break;
                nc[big] = nc[big] * 3 / 4 < nc_min ? nc_min : nc[big] * 3 / 4;
            }
            for (int k = 0; k < 3; ++k) ws->ncell[k] = nc[k];
            ca = CellArg{nc[0], nc[1], nc[2], nc[0] / box[0], nc[1] / box[1], nc[2] / box[2], stencil};
        }
    }
    // Verlet skin: decide on the device whether this call rebuilds (first call with these atoms / box / cutoff: always)
    const int32_t* guard = nullptr;
    bool forced = false;
    if (use_skin) {
        pantea_workspace::SkinKey key;
        key.n = n; key.own_begin = ws->own_begin; key.own_end = ws->own_end; key.rc = rc; key.types = types;
        for (int k = 0; k < 3; ++k) key.box[k] = box[k];
        if (ws->skin_active && key == ws->skin_key) {
            const double half = 0.5 * ws->skin;
            skin_check_kernel<T><<<blocks_n, threads, 0, st>>>(pos, (const T*)ws->pos_ref, (int)n, ba, half * half, ws->skin_flags);
            PANTEA_LAUNCH_CHECK();
        } else {
            PANTEA_CUDA_TRY(cudaMemsetAsync(ws->skin_flags, 1, 4, st));  // flags[0] != 0: rebuild
            ws->lists_valid = false;
            forced = true;
        }
        ws->skin_key = key;
        guard = ws->skin_flags;
    }
    ws->skin_active = use_skin;
    if (!use_skin) ws->lists_valid = false;
    const int own_lo = (int)ws->own_begin, own_hi = ws->own_end < 0 ? (int)n : (int)ws->own_end;
    const uint8_t* role = ws->role;
    if (role && !use_cells) return fail(PANTEA_EINVAL, "brick decomposition needs a box of at least three cells per axis");
    const bool owned = use_cells && (role != nullptr || own_lo > 0 || own_hi < (int)n);
    ws->owned_active = owned;
    if (use_cells) {
        const int64_t ncells = (int64_t)ca.nx * ca.ny * ca.nz;
        int rcode = ensure_cell_capacity(ws, ncells);
        if (rcode != PANTEA_OK) return rcode;
        ws->mode = kModeCell;
        const bool was_clean = ws->scratch_clean;
        if (!guard || forced) {  // device-decided rebuilds find the scratch zeroed by the previous rebuild's kernels
            // every completed build leaves the counting-sort scratch zeroed (cell_sort_pack / the scan do it), so only
            // the first build after an allocation -- or after a build that failed half-way -- has to clear it
            if (!ws->scratch_clean) {
                PANTEA_CUDA_TRY(cudaMemsetAsync(ws->cell_fill, 0, 4 * (ws->cell_cap + 1), st));
                PANTEA_CUDA_TRY(cudaMemsetAsync(ws->cell_own_cnt, 0, 4 * (ws->cell_cap + 1), st));
            }
            PANTEA_CUDA_TRY(cudaMemsetAsync(ws->wide_flag, 0, 8, st));
        }
        (void)was_clean;
        ws->scratch_clean = false;  // until this build's kernels are all queued
        cell_assign_kernel<T><<<blocks_n, threads, 0, st>>>(pos, (int)n, ca, ba, ws->cell_of, ws->cell_fill, ws->wide_flag, guard, role,
                                                            owned ? ws->cell_own_cnt : nullptr, own_lo, own_hi);
        PANTEA_LAUNCH_CHECK();
        rcode = launch_cell_scan(ws, ws->cell_fill, (int)ncells, ws->cell_start, ws->cell_fill,
                                 owned ? ws->cell_own_cnt : nullptr, ws->cell_own, ws->cell_own_cnt, guard, st);
        if (rcode != PANTEA_OK) return rcode;
        cell_scatter_kernel<<<blocks_n, threads, 0, st>>>(ws->cell_of, (int)n, ws->cell_start, ws->cell_fill, ws->tmp_order, guard);
        PANTEA_LAUNCH_CHECK();
        const int blocks_c = (int)((ncells * 32 + threads - 1) / threads);
        cell_sort_pack_kernel<T><<<blocks_c, threads, 0, st>>>(pos, types, ws->n_types, ws->cell_start, (int)ncells,
                                                               ws->tmp_order, rec, ws->slot_of,
                                                               (Rec<float>*)ws->rec_screen, ba, ws->cell_fill, guard,
                                                               own_lo, own_hi, owned ? ws->cell_own : nullptr, ws->nbr_tcount, role, ws->wide_flag + 1);
        PANTEA_LAUNCH_CHECK();
        if (owned) {  // compact list of the owned atoms' slots (cell order) behind the scanned per-cell owned counts
            owned_fill_kernel<T><<<blocks_c, threads, 0, st>>>(rec, ws->cell_start, (int)ncells, ws->cell_own, own_lo, own_hi,
                                                               ws->owned_slots, guard, role);
            PANTEA_LAUNCH_CHECK();
        }
    } else {
        ws->mode = kModeAllPairs;
        pack_identity_kernel<T><<<blocks_n, threads, 0, st>>>(pos, types, ws->n_types, (int)n, struct_ptr, (int)n_structs,
                                                              rec, ws->slot_of, ws->struct_of);
        PANTEA_LAUNCH_CHECK();
    }
    RowArgs ra;
    ra.n = (int)n; ra.cap = ws->cap; ra.n_buckets = ws->n_types + 1;
    ra.box = ba; ra.cell = ca; ra.cell_start = ws->cell_start; ra.struct_of = ws->struct_of;
    ra.struct_ptr = struct_ptr; ra.boxes = boxes; ra.rc = rc;
    ra.own_begin = (int)ws->own_begin; ra.own_end = ws->own_end < 0 ? (int)n : (int)ws->own_end;
    ra.nbr = ws->nbr; ra.tcount = ws->nbr_tcount; ra.flags = ws->flags;
    ra.rec_screen = use_cells ? (const Rec<float>*)ws->rec_screen : nullptr;
    ra.wide_flag = ws->wide_flag;
    ra.guard = guard;
    ra.owned_slots = owned ? ws->owned_slots : nullptr;
    ra.n_owned = role ? (int)ws->own_cap : own_hi - own_lo;
    ra.n_owned_dev = role ? ws->cell_own + (int64_t)ca.nx * ca.ny * ca.nz : nullptr;
    ra.role = role;
    {   // error bound of the screening distance (see neighbor_rows_kernel): per-axis |error| <= 5 L 2^-24, squared
        // distance |error| <= 2 sqrt(3) r delta + 3 delta^2 + 4 2^-24 r^2 at r ~ rc; doubled for safety
        const double lmax = std::max(ws->box[0], std::max(ws->box[1], ws->box[2]));
        const double delta = 5.0 * lmax * 5.9604644775390625e-08;
        ra.screen_band = (float)(2.0 * (3.5 * rc * delta + 3.0 * delta * delta + 3.0e-7 * rc * rc));
    }
    const int blocks_w = (int)(((owned ? (int64_t)ra.n_owned : n) + kWarpsPerBlock - 1) / kWarpsPerBlock);
    // scratch of the row-per-lane scan: 125 half-width cells hold ~3.5 cap candidates on average
    ra.tmp_cap = (use_cells && ca.r == 2 && ra.rec_screen) ? ((6 * ws->cap + 31) / 32 * 32) : 0;
    const size_t smem = (size_t)kWarpsPerBlock * (ws->cap + ra.tmp_cap) * sizeof(int32_t);
    {
        static size_t configured[64] = {0};
        int rc_s = use_cells ? opt_in_smem((const void*)neighbor_rows_kernel<T, kModeCell>, smem, configured, "neighbour rows: capacity too large for shared memory")
                             : PANTEA_OK;
        if (rc_s != PANTEA_OK) return rc_s;
    }
    if (use_cells)
        neighbor_rows_kernel<T, kModeCell><<<blocks_w, kWarpsPerBlock * 32, smem, st>>>(rec, ra);
    else
        neighbor_rows_kernel<T, kModeAllPairs><<<blocks_w, kWarpsPerBlock * 32, smem, st>>>(rec, ra);
    PANTEA_LAUNCH_CHECK();
    if (use_cells) ws->scratch_clean = true;
    if (use_skin) {
        skin_refresh_kernel<T><<<blocks_n, threads, 0, st>>>(pos, rec, (int)n);
        PANTEA_LAUNCH_CHECK();
        skin_save_kernel<T><<<(int)((3 * n + threads - 1) / threads), threads, 0, st>>>(pos, (T*)ws->pos_ref, (int)(3 * n), guard);
        PANTEA_LAUNCH_CHECK();
        skin_finish_kernel<<<1, 1, 0, st>>>(ws->skin_flags, ws->wide_flag);
        PANTEA_LAUNCH_CHECK();
    }
    return PANTEA_OK;
}

int neighbor_build_impl(pantea_workspace* ws, const void* pos, const int32_t* types, int64_t n, const double* box,
                        const int32_t* struct_ptr, const double* boxes, int64_t n_structs, double rc, cudaStream_t st) {
    if (!ws || !pos || !types) return fail(PANTEA_EINVAL, "pantea_neighbor_build: NULL argument");
    if (n < 0 || n > ws->max_atoms) return fail(PANTEA_EINVAL, "pantea_neighbor_build: n_atoms exceeds the workspace capacity");
    if (!(rc > 0.0)) return fail(PANTEA_EINVAL, "pantea_neighbor_build: r_cutoff must be positive");
    if (box && !(box[0] > 0.0 && box[1] > 0.0 && box[2] > 0.0)) return fail(PANTEA_EINVAL, "pantea_neighbor_build: box lengths must be positive");
    if (n == 0) { ws->n = 0; ws->mode = kModeAllPairs; return PANTEA_OK; }
    if (ws->dtype == PANTEA_F64) return build_typed<double>(ws, pos, types, n, box, struct_ptr, boxes, n_structs, rc, st);
    return build_typed<float>(ws, pos, types, n, box, struct_ptr, boxes, n_structs, rc, st);
}

}  // namespace pantea

using namespace pantea;

extern "C" {

int pantea_neighbor_build(pantea_workspace* ws, const void* positions, const int32_t* types, int64_t n_atoms,
                          const double* box, double r_cutoff, void* stream) {
    return neighbor_build_impl(ws, positions, types, n_atoms, box, nullptr, nullptr, 1, r_cutoff, (cudaStream_t)stream);
}

int pantea_neighbor_build_batch(pantea_workspace* ws, const void* positions, const int32_t* types, int64_t n_atoms,
                                const int32_t* struct_ptr, const double* boxes, int64_t n_structs, double r_cutoff,
                                void* stream) {
    if (!struct_ptr || n_structs < 1) return fail(PANTEA_EINVAL, "pantea_neighbor_build_batch: struct_ptr/n_structs invalid");
    return neighbor_build_impl(ws, positions, types, n_atoms, nullptr, struct_ptr, boxes, n_structs, r_cutoff,
                               (cudaStream_t)stream);
}

int pantea_neighbor_status(pantea_workspace* ws, int32_t* max_count, void* stream) {
    if (!ws) return fail(PANTEA_EINVAL, "pantea_neighbor_status: NULL workspace");
    int32_t h[4] = {0, 0, 0, 0};
    PANTEA_CUDA_TRY(cudaMemcpyAsync(h, ws->flags, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PANTEA_CUDA_TRY(cudaMemsetAsync(ws->flags, 0, 16, (cudaStream_t)stream));  // maxima are sticky until read
    PANTEA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (max_count) *max_count = h[0];
    int code = PANTEA_OK;
    std::string msg;
    if (h[1] > 0) {
        code = PANTEA_ECAPACITY;
        msg = "neighbour block overflow in the evaluation kernels: " + std::to_string(h[1]) + " neighbours > staged capacity " +
              std::to_string(ws->smem_cap) + " (re-run: the capacity has been raised)";
    }
    const int seen = h[0] > h[1] ? h[0] : h[1];  // maxima since the last status call (0: nothing was built since)
    if (seen > 0 && h[0] <= ws->cap) {
        // shared-memory footprint of the evaluation kernels follows the observed maximum (+10 % + 8 head-room); the fast
        // path keeps four blocks of four 80-byte-record neighbour blocks per SM only up to 144 staged neighbours, so
        // the head-room gives way down to 8 there
        int want = (seen + seen / 10 + 8 + 7) / 8 * 8;
        if (ws->pot && ws->pot->v2_ok && want > 144 && seen + 8 <= 144) want = 144;
        if (want < 32) want = 32;
        if (want > ws->cap) want = ws->cap;
        // sized from the first observation, afterwards it only grows: shrinking on a later, smaller observation would
        // oscillate when the same trajectory (e.g. a densifying box) is run again after a capacity report
        if (ws->smem_cap == 0 || want > ws->smem_cap) {
            ws->smem_cap = want;
            ws->lists_valid = false;  // pair lists were cut to the old staged capacity
            ++ws->arg_epoch;
        }
    }
    if (h[2] > 0) {  // pair lists: capacity follows the observed maximum (+25 % head-room)
        if (ws->pair_cap > 0 && h[2] > ws->pair_cap) {
            code = PANTEA_ECAPACITY;
            msg = "pair-list overflow: " + std::to_string(h[2]) + " pairs > capacity " + std::to_string(ws->pair_cap) +
                  " (re-run: the capacity has been raised)";
        }
        const int want = h[2] + h[2] / 4 + 64;
        // grow whenever needed; shrink only once, when the first observation shows the initial guess to be far too large
        if (want > ws->pair_cap || (ws->pair_cap_request == 0 && want < ws->pair_cap / 2)) { ws->pair_cap_request = want; ++ws->arg_epoch; }
    }
    if (code != PANTEA_OK && h[0] <= ws->cap) return fail(code, msg);
    if (h[0] > ws->cap)
        return fail(PANTEA_ECAPACITY, "neighbour row overflow: " + std::to_string(h[0]) + " neighbours > capacity " +
                                          std::to_string(ws->cap));
    return PANTEA_OK;
}

int pantea_neighbor_counts(pantea_workspace* ws, int32_t* counts, void* stream) {
    if (!ws || !counts) return fail(PANTEA_EINVAL, "pantea_neighbor_counts: NULL argument");
    if (ws->mode == kModeNone) return fail(PANTEA_EINVAL, "pantea_neighbor_counts: no structure bound");
    if (ws->skin_active) return fail(PANTEA_EINVAL, "pantea_neighbor_counts: rows were built with a Verlet skin (set skin = 0)");
    if (ws->n == 0) return PANTEA_OK;
    neighbor_counts_kernel<<<(int)((ws->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ws->slot_of, ws->nbr_tcount,
                                                                                       (int)ws->n, counts);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int pantea_neighbor_export(pantea_workspace* ws, const int64_t* row_ptr, int32_t* col_idx, void* stream) {
    if (!ws || !row_ptr || !col_idx) return fail(PANTEA_EINVAL, "pantea_neighbor_export: NULL argument");
    if (ws->mode == kModeNone) return fail(PANTEA_EINVAL, "pantea_neighbor_export: no structure bound");
    if (ws->skin_active) return fail(PANTEA_EINVAL, "pantea_neighbor_export: rows were built with a Verlet skin (set skin = 0)");
    if (ws->n == 0) return PANTEA_OK;
    const int blocks = (int)((ws->n + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const size_t smem = (size_t)kWarpsPerBlock * ws->cap * 4;
    if (ws->dtype == PANTEA_F64)
        neighbor_export_kernel<double><<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
            (const Rec<double>*)ws->rec, ws->slot_of, ws->nbr_tcount, ws->nbr, ws->cap, (int)ws->n, row_ptr, col_idx);
    else
        neighbor_export_kernel<float><<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
            (const Rec<float>*)ws->rec, ws->slot_of, ws->nbr_tcount, ws->nbr, ws->cap, (int)ws->n, row_ptr, col_idx);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int pantea_distances(pantea_workspace* ws, const int32_t* idx_i, int64_t n_i, const int32_t* idx_j, int64_t n_j, void* r,
                     void* d, void* stream) {
    if (!ws || !r) return fail(PANTEA_EINVAL, "pantea_distances: NULL argument");
    if (ws->mode == kModeNone) return fail(PANTEA_EINVAL, "pantea_distances: no structure bound");
    if (!idx_i) n_i = ws->n;
    if (!idx_j) n_j = ws->n;
    if (n_i * n_j == 0) return PANTEA_OK;
    BoxArg ba{ws->box[0], ws->box[1], ws->box[2], ws->has_box ? 1 : 0};
    const int64_t total = n_i * n_j;
    const int blocks = (int)((total + 255) / 256);
    if (ws->dtype == PANTEA_F64)
        distances_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const Rec<double>*)ws->rec, ws->slot_of,
            ws->struct_of, ws->boxes, ba, idx_i, n_i, idx_j, n_j, (double*)r, (double*)d);
    else
        distances_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const Rec<float>*)ws->rec, ws->slot_of,
            ws->struct_of, ws->boxes, ba, idx_i, n_i, idx_j, n_j, (float*)r, (float*)d);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

}  // extern "C"
