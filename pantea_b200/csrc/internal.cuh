// Internal declarations shared by the translation units of libpantea_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "pantea_b200.h"

namespace pantea {

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
void count_launch(int n = 1);
int64_t launch_count();

#define PANTEA_CUDA_TRY(expr)                                                                         \
    do {                                                                                              \
        cudaError_t err__ = (expr);                                                                   \
        if (err__ != cudaSuccess)                                                                     \
            return ::pantea::fail(PANTEA_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    } while (0)

#define PANTEA_LAUNCH_CHECK()                                                                         \
    do {                                                                                              \
        cudaError_t err__ = cudaGetLastError();                                                       \
        if (err__ != cudaSuccess)                                                                     \
            return ::pantea::fail(PANTEA_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(err__)); \
        ::pantea::count_launch();                                                                     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side potential tables (parameters kept in double; kernels convert to their arithmetic type)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTypes = PANTEA_MAX_TYPES;
constexpr int kMaxSF = PANTEA_MAX_SYMFUNC;
constexpr int kMaxLayers = PANTEA_MAX_LAYERS;
constexpr int kMaxCut = PANTEA_MAX_CUTOFFS;
constexpr int kMaxGroups = 64;
constexpr int kBuckets = kMaxTypes + 1;  // neighbour rows are partitioned by type; last bucket = "other"

struct RadialSF {
    int type_j;  // 0-based bucket
    int kind;    // 1 | 2
    int cls;     // cutoff class
    int out;     // output column
    double eta, r_shift;
};

struct AngularMember {
    double eta, lambda0, zeta, pref;  // pref = 2^(1-zeta)
    int izeta;                        // zeta when it is a small non-negative integer, else -1
    int out;
};

struct AngularGroup {
    int type_j, type_k;  // 0-based buckets
    int cls;
    int kind;  // 3 | 9
    int first, count;  // members [first, first+count)
};

struct CutoffClass {
    int type;
    double rc;
};

struct ElementTable {
    int n_sf, n_radial, n_groups, n_cls;
    double rc_max;
    CutoffClass cls[kMaxCut];
    RadialSF radial[kMaxSF];
    AngularGroup groups[kMaxGroups];
    AngularMember members[kMaxSF];
    int has_scaler;
    double shift[kMaxSF], slope[kMaxSF], offset[kMaxSF];
    int n_layers;
    int sizes[kMaxLayers + 1];
    int acts[kMaxLayers];
    int w_off[kMaxLayers];  // offset of the layer's kernel in `weights` (bias follows the kernel)
    int n_neurons;          // sum of sizes[1..]
    const double* weights;  // device
    // fast path of acsf2.cu: one cutoff class (tanhu), every angular group a single G3 member with integer zeta >= 1,
    // every neighbour type in at most one group -> one (W, Q) weight pair per staged neighbour
    int v2_ok;
    int v2_group_of_type[kBuckets];  // the group a neighbour of this type takes part in, -1: none
    double v2_eta[kBuckets], v2_wscale[kBuckets];  // that group's eta and sqrt(2^(1-zeta)) (0: no group)
};

}  // namespace pantea

struct pantea_potential {
    int n_elements = 0;
    double rc_max = 0.0;
    int max_sf = 0, max_cls = 1, max_neurons = 0, max_width = 0, max_members = 1, max_groups = 0;
    bool v2_ok = true;  // every element table qualifies for the fast path of acsf2.cu
    std::vector<pantea::ElementTable> host;
    pantea::ElementTable* dev = nullptr;  // [n_elements]
    std::vector<double*> dev_weights;
};

namespace pantea {

// packed atom record in cell-sorted order
template <typename T>
struct Rec;
template <>
struct __align__(16) Rec<double> {
    double x, y, z;
    int type;  // 0-based bucket
    int idx;   // original atom index
};
template <>
struct __align__(16) Rec<float> {
    float x, y, z;
    int tidx;  // bucket << 28 | original index
};

template <typename T>
__host__ __device__ inline int rec_type(const Rec<T>& r);
template <>
__host__ __device__ inline int rec_type<double>(const Rec<double>& r) { return r.type; }
template <>
__host__ __device__ inline int rec_type<float>(const Rec<float>& r) { return (int)((unsigned)r.tidx >> 28); }
template <typename T>
__host__ __device__ inline int rec_idx(const Rec<T>& r);
template <>
__host__ __device__ inline int rec_idx<double>(const Rec<double>& r) { return r.idx; }
template <>
__host__ __device__ inline int rec_idx<float>(const Rec<float>& r) { return r.tidx & 0x0fffffff; }
__host__ __device__ inline void rec_set(Rec<double>& r, int type, int idx) { r.type = type; r.idx = idx; }
__host__ __device__ inline void rec_set(Rec<float>& r, int type, int idx) { r.tidx = (int)(((unsigned)type << 28) | (unsigned)idx); }

enum NeighborMode { kModeNone = 0, kModeCell = 1, kModeAllPairs = 2 };

}  // namespace pantea

struct pantea_workspace {
    const pantea_potential* pot = nullptr;
    int64_t max_atoms = 0;
    int cap = 0;    // neighbours per row
    void* gbuf = nullptr;          // [max_atoms][max_sf][4] summed descriptors (evaluation -> network kernel)
    void* wbuf = nullptr;          // [max_atoms][max_sf] dE_i/dG_is (network kernel -> full-force scatter pass)
    int32_t* pairs = nullptr;      // [max_atoms][pair_cap] pre-filtered (j,k) pair lists
    int32_t* pair_off = nullptr;   // [max_atoms][pair_groups + 1]
    int pair_cap = 0, pair_groups = 0, pair_cap_request = 0;
    int smem_cap = 0;  // rows staged in shared memory by the atom kernel (0: cap); set from the observed maximum
    int dtype = PANTEA_F64;
    bool compute32 = false;  // mixed mode: symmetry functions in single precision on double-precision state (fast path only)
    int n_types = 0;  // buckets used by the potential (others -> bucket n_types)

    // current binding
    int64_t n = 0;
    int mode = pantea::kModeNone;
    bool has_box = false;
    double box[3] = {0, 0, 0};
    double rc = 0.0;
    int ncell[3] = {1, 1, 1};
    int64_t n_structs = 0;
    const int32_t* struct_ptr = nullptr;  // borrowed (batch mode)
    const double* boxes = nullptr;        // borrowed (batch mode)
    int64_t own_begin = 0, own_end = -1;  // -1: all atoms
    // brick-decomposed multi-GPU engine (mgpu.cu): per-atom role instead of an index range -- 0 absent on this rank
    // (not binned), 1 ghost, 2 owned; the owned count is then only known on the device (cell_own[ncells])
    const uint8_t* role = nullptr;  // borrowed, [n]
    int64_t own_cap = 0;            // static upper bound of the owned atoms (launch grids)

    // device buffers
    void* rec = nullptr;           // Rec<T>[max_atoms]
    void* rec_screen = nullptr;    // Rec<float>[max_atoms]: box-wrapped positions for the FP32 screening pass (F64 cell mode)
    int32_t* wide_flag = nullptr;  // [1] set when an atom lies more than a quarter box outside the cell
    // Verlet skin (pantea_workspace_set_skin)
    double skin = 0.0;
    void* pos_ref = nullptr;        // [max_atoms,3] positions at the last rebuild (workspace dtype)
    int32_t* skin_flags = nullptr;  // device [4]: [0] rebuild in this call, [1] pair lists stale, [2] rebuilds, [3] skin builds
    bool skin_active = false;       // rows currently hold radius rc + skin and pos_ref matches them
    bool lists_valid = false;       // pair lists belong to the energy pass over the current rows
    bool lists_v2 = false;          // ... and are in the fast path's format (acsf2.cu)
    struct SkinKey {
        int64_t n = -1, own_begin = 0, own_end = -1;
        double rc = 0, box[3] = {0, 0, 0};
        const void* types = nullptr;
        bool operator==(const SkinKey& o) const {
            return n == o.n && own_begin == o.own_begin && own_end == o.own_end && rc == o.rc && box[0] == o.box[0] &&
                   box[1] == o.box[1] && box[2] == o.box[2] && types == o.types;
        }
    } skin_key;
    int32_t* slot_of = nullptr;    // [max_atoms] original index -> sorted slot
    int32_t* struct_of = nullptr;  // [max_atoms] structure id per sorted slot (batch mode)
    int32_t* nbr = nullptr;        // [max_atoms * cap] sorted-slot indices, partitioned by bucket
    int32_t* nbr_tcount = nullptr; // [max_atoms * kBuckets]
    int32_t* cell_of = nullptr;    // [max_atoms]
    int32_t* tmp_order = nullptr;  // [max_atoms]
    int32_t* cell_start = nullptr; // [cell_cap + 1]
    int32_t* cell_fill = nullptr;  // [cell_cap]
    int32_t* scan_sums = nullptr;  // [2 * (cell_cap / 2048 + 2)] tile sums of the two-launch cell scan
    int32_t* cell_own = nullptr;   // [cell_cap + 1] exclusive scan of the owned atoms per cell (ranks that own a part)
    int32_t* cell_own_cnt = nullptr;  // [cell_cap + 1] owned atoms per cell (counted while binning, zeroed by the scan)
    int32_t* owned_slots = nullptr; // [max_atoms] cell-ordered slots of the owned atoms, ascending
    bool owned_active = false;     // rows / evaluation run over `owned_slots` (cell mode with a proper owned range)
    int64_t cell_cap = 0;
    bool scratch_clean = false;    // cell_fill / cell_own_cnt are all zero (left so by every completed build)
    int32_t* flags = nullptr;      // [4]: max neighbour count seen, ...
    double* e_partial = nullptr;   // reduction scratch
    int64_t e_partial_cap = 0;
    void* md_forces = nullptr;     // [max_atoms,3] F(t+dt) scratch for pantea_md_run
    double* md_ke = nullptr;       // [4]
    unsigned long long* counters = nullptr;  // optional device work counters [4] (borrowed)
    void* md_eatom = nullptr;      // [max_atoms]
    cudaGraphExec_t md_graph = nullptr;
    cudaStream_t capture_stream = nullptr;
    int md_graph_nodes = 0;        // kernel launches per replay
    int64_t arg_epoch = 0;         // bumped whenever a buffer / capacity baked into captured kernel arguments changes
    // key of the captured graph
    struct GraphKey {
        const void *pos = nullptr, *vel = nullptr, *frc = nullptr, *mass = nullptr, *types = nullptr, *scalars = nullptr;
        int64_t n = 0, epoch = -1;
        double dt = 0, tau = 0, t0 = 0, kb = 0, box[3] = {0, 0, 0};
        int record = 0, has_box = 0, mass_scaled = 0, force_mode = 0;
        bool operator==(const GraphKey& o) const {
            return pos == o.pos && vel == o.vel && frc == o.frc && mass == o.mass && types == o.types &&
                   scalars == o.scalars && n == o.n && dt == o.dt && tau == o.tau && t0 == o.t0 && kb == o.kb &&
                   box[0] == o.box[0] && box[1] == o.box[1] && box[2] == o.box[2] && record == o.record &&
                   has_box == o.has_box && epoch == o.epoch && mass_scaled == o.mass_scaled && force_mode == o.force_mode;
        }
    } md_key;
};

namespace pantea {
// implemented in neighbor.cu
int neighbor_build_impl(pantea_workspace* ws, const void* pos, const int32_t* types, int64_t n, const double* box,
                        const int32_t* struct_ptr, const double* boxes, int64_t n_structs, double rc, cudaStream_t st);
// implemented in acsf.cu
int atom_kernel_launch(pantea_workspace* ws, int element_slot, const int32_t* centres, int64_t n_centres, void* G,
                       void* dG, void* e_atom, void* forces, cudaStream_t st, int force_mode = PANTEA_FORCE_REFERENCE);
int reduce_energy(pantea_workspace* ws, const void* e_atom, void* e_total, cudaStream_t st);
// implemented in acsf2.cu: pair filter + evaluation of the fast path (double precision, gradients, no r_jk wrap)
template <typename T> struct AtomArgs;
int launch_v2(const AtomArgs<double>& a, cudaStream_t st);
int ensure_cell_capacity(pantea_workspace* ws, int64_t ncells);
int opt_in_smem(const void* kern, size_t smem, size_t* configured /*[64], per device*/, const char* what);
}  // namespace pantea
