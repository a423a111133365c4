// Definitions shared by the symmetry-function translation units (acsf.cu: generic kernels; acsf2.cu: the
// specialised single-class / single-member fast path).
#pragma once
#include "internal.cuh"
#include "math.cuh"

#ifndef PANTEA_EVAL_MINBLOCKS
#define PANTEA_EVAL_MINBLOCKS 4  // resident 128-thread blocks per SM the FP64 evaluation kernel is compiled for
#endif
#ifndef PANTEA_TRIPLETS_PER_LANE
#define PANTEA_TRIPLETS_PER_LANE 1
#endif

namespace pantea {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kEvalWarps = 4;       // eval kernel: 4 warps per block, shared by 4 / WPA atoms (WPA = warps per atom)
#ifndef PANTEA_EVAL_WPA
#define PANTEA_EVAL_WPA 1           // warps per atom when there are enough atoms to fill the GPU
#endif
constexpr int kFilterWarps = 8;     // pair filter: 8 warps per block, one atom each
constexpr int kNU = PANTEA_TRIPLETS_PER_LANE;
#ifndef PANTEA_STAGE_ITERS
#define PANTEA_STAGE_ITERS 8
#endif
constexpr int kStageIters = PANTEA_STAGE_ITERS;  // pair-list iterations staged per cp.async group

struct BoxArgK {
    double lx, ly, lz;
    int has_box;
};

template <typename T>
struct AtomArgs {
    const Rec<T>* rec;
    const int32_t* nbr;
    const int32_t* tcount;
    int cap;   // row stride of `nbr`
    int scap;  // neighbour records staged in shared memory (<= cap; longer rows raise the overflow flag)
    int32_t* flags;
    const int32_t* slot_of;
    const int32_t* struct_of;
    const double* boxes;
    BoxArgK box;
    int wrap_jk;
    double rc_list;  // radius the neighbour rows were built with (cutoff + Verlet skin)
    float skin;      // Verlet skin: the pair lists must stay valid while atoms move by up to skin / 2 each
    const int32_t* filter_guard;  // skin: the filter is skipped while *filter_guard == 0 (NULL: always run)
    int apw;                  // fast path: atoms per warp of an evaluation block (drawn from a shared counter)
    int rec_bytes;            // fast path: bytes per staged neighbour record (80: double evaluation, 48: single)
    const int32_t* dup_flag;  // != 0 when two present atoms may coincide (cell-list binning): the pair filters then test exactly
    int dup_always;           // all-pairs mode has no such flag: always test
    float screen_t;  // fast path: triplets whose Gaussian weight is below exp(-screen_t) of the group's largest are skipped (0: off)
    const ElementTable* tables;
    int n_types;
    int element_slot;  // >= 0: apply this element's table to every centre; -1: the atom's own type
    const int32_t* centres;
    int n_work;
    int by_slot;  // 1: work item = cell-sorted slot (energy/force pass); 0: work item = centre list entry
    int own_begin, own_end;
    const int32_t* owned_slots;  // by_slot passes of a block-owned rank: work item -> slot (NULL: work item = slot)
    const int32_t* n_work_dev;   // role mode (brick decomposition): number of work items on the device, n_work is its bound
    // pair lists written by the filter, read by the evaluation
    int32_t* pairs;      // [n_work][pair_cap]  (j | k << 16), row positions within the staged neighbour block
    int32_t* pair_off;   // [n_work][max_groups + 1] offsets of each group's segment
    int pair_cap, max_groups;
    T* G;
    T* dG;
    int g_stride;
    T* e_atom;
    T* forces;
    T* gbuf;  // [n_work][n_sf_max][4] summed descriptors handed from the evaluation to the network kernel
    T* wbuf;  // full-force mode: [n_work][n_sf_max] dE_i/dG_is written by the network kernel, read by the scatter pass
    unsigned long long* counters;  // optional work counters: [0] pairs, [1] radial-SF evals, [2] triplet-SF evals
    int n_cls_max, n_sf_max, n_neurons_max, width_max;
};

template <typename T>
__host__ __device__ inline size_t eval_smem_bytes(int cap, int n_cls, int n_sf, int n_neurons, int width, int wpa) {
    size_t t_elems = (size_t)(5 + 2 * n_cls) * (cap + 1)  // neighbour records + one all-zero padding record
                     + (size_t)wpa * n_sf * 4;            // per-warp partial sums
    (void)n_neurons; (void)width;
    return (t_elems * sizeof(T) + 15) & ~size_t(15);
}

template <typename T>
__device__ __forceinline__ T powi(T base, int n) {
    T r = (T)1;
    while (n > 0) {
        if (n & 1) r *= base;
        base *= base;
        n >>= 1;
    }
    return r;
}

// rarely used, large library routines are kept out of line so that the hot loops stay small
template <typename T>
__device__ __noinline__ T pow_general(T base, T e) { return t_pow<T>(base, e); }
template <typename T>
__device__ __noinline__ void cutoff_eval_ool(int type, T r, T rc, T* fc, T* dfc) {
    T a, b;
    cutoff_eval<T>(type, r, rc, a, b);
    *fc = a; *dfc = b;
}
template <typename T>
__device__ __noinline__ void activation_eval_ool(int act, T x, T* y, T* dy) {
    T a, b;
    activation_eval<T>(act, x, a, b);
    *y = a; *dy = b;
}

// barrier over the WPA warps of one atom (several atoms share a block: named barriers 1.., one per atom)
template <int WPA>
__device__ __forceinline__ void group_sync(int atom_in_block) {
    if (WPA == 1) __syncwarp();
    else if (WPA == kEvalWarps) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(atom_in_block + 1), "n"(WPA * 32) : "memory");
}

// work item -> (cell-ordered slot, output row, element table); false when the item is not evaluated
template <typename T>
__device__ __forceinline__ bool resolve_item(const AtomArgs<T>& a, int w, int& slot, int& out_row, int& etype) {
    if (a.by_slot) {
        if (a.n_work_dev && w >= *a.n_work_dev) return false;
        slot = a.owned_slots ? a.owned_slots[w] : w;
        out_row = rec_idx(a.rec[slot]);
        if (!a.n_work_dev && (out_row < a.own_begin || out_row >= a.own_end)) return false;
    } else {
        const int oi = a.centres ? a.centres[w] : w;
        slot = a.slot_of[oi];
        out_row = w;
    }
    etype = a.element_slot >= 0 ? a.element_slot : rec_type(a.rec[slot]);
    return true;
}

template <typename T>
__device__ __forceinline__ void item_box(const AtomArgs<T>& a, int slot, T& lx, T& ly, T& lz, bool& pbc) {
    lx = (T)a.box.lx; ly = (T)a.box.ly; lz = (T)a.box.lz;
    pbc = a.box.has_box != 0;
    if (a.boxes) {
        const int s = a.struct_of[slot];
        lx = (T)a.boxes[3 * s]; ly = (T)a.boxes[3 * s + 1]; lz = (T)a.boxes[3 * s + 2];
        pbc = true;
    }
}

// neighbour segments by type bucket: seg[b] .. seg[b+1] within the (type-partitioned) row, clamped to `total`
struct Segments {
    int seg[kBuckets + 1];
    int total;
    __device__ __forceinline__ void load(const int32_t* tc, int cap) {
        int acc = 0;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) { seg[b] = acc; acc += tc[b]; }
        seg[kBuckets] = acc;
        total = acc < cap ? acc : cap;
    }
    __device__ __forceinline__ int lo(int t) const {
        int v = 0;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) if (b == t) v = seg[b];
        return v < total ? v : total;
    }
    __device__ __forceinline__ int hi(int t) const {
        int v = 0;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) if (b == t) v = seg[b + 1];
        return v < total ? v : total;
    }
};

}  // namespace pantea
