/*
 * pantea_b200 -- C ABI of the B200-native HDNNP energy/force hot path.
 *
 * The reference (hghcomphys/pantea, Python/JAX) has no FFI layer: its "operator interface" for
 * this path is the set of module-level jitted kernels that take plain arrays.  Every entry
 * point below replaces one of them (reference file:line given per function, paths relative to
 * the pantea repository) and is what a binding on the reference side (ctypes, or an XLA FFI
 * custom-call wrapping the same symbols -- see INTEGRATION.md) would bind.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Unless stated otherwise pointers are DEVICE pointers
 *     (borrowed, never freed or resized here); descriptors (`*_desc`) and `box` are HOST memory.
 *   - `dtype` is PANTEA_F64 or PANTEA_F32 and selects the arithmetic and the element type of
 *     every `void*` array of that workspace (reference: `default_dtype.FLOATX`, types.py:13-30).
 *   - Atom types are 1-based and ordered by ascending atomic number among the potential's
 *     elements (reference element.py:99-108).  Types outside 1..n_elements are legal atoms that
 *     no symmetry function refers to.
 *   - All functions are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant per
 *     workspace handle, and return 0 on success or a negative PANTEA_E* code; the message is
 *     available from pantea_last_error() (thread local).  Device-side capacity overflow is
 *     reported by pantea_neighbor_status().
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     PANTEA_ECUDA.
 */
#ifndef PANTEA_B200_H
#define PANTEA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PANTEA_OK 0
#define PANTEA_EINVAL (-1)
#define PANTEA_ECUDA (-2)
#define PANTEA_ECAPACITY (-3)
#define PANTEA_ENOMEM (-4)

#define PANTEA_F64 64
#define PANTEA_F32 32

#define PANTEA_MAX_TYPES 8       /* elements per potential */
#define PANTEA_MAX_SYMFUNC 128   /* symmetry functions per element */
#define PANTEA_MAX_LAYERS 8      /* dense layers per element network */
#define PANTEA_MAX_CUTOFFS 4     /* distinct (cutoff type, radius) pairs per element */

/* symmetry function kinds (RuNNer numbering; reference potential.py:213-252) */
#define PANTEA_G1 1
#define PANTEA_G2 2
#define PANTEA_G3 3
#define PANTEA_G9 9

/* cutoff function types (RuNNer `cutoff_type`; reference settings.py:43-51, cutoff.py:45-110) */
enum { PANTEA_CUT_HARD = 0, PANTEA_CUT_COS = 1, PANTEA_CUT_TANHU = 2, PANTEA_CUT_TANH = 3,
       PANTEA_CUT_EXP = 4, PANTEA_CUT_POLY1 = 5, PANTEA_CUT_POLY2 = 6 };

/* activation functions (reference activation.py:48-60; EXP is exp(-x)) */
enum { PANTEA_ACT_IDENTITY = 0, PANTEA_ACT_TANH = 1, PANTEA_ACT_LOGISTIC = 2, PANTEA_ACT_SOFTPLUS = 3,
       PANTEA_ACT_RELU = 4, PANTEA_ACT_GAUSSIAN = 5, PANTEA_ACT_COS = 6, PANTEA_ACT_EXP = 7,
       PANTEA_ACT_HARMONIC = 8 };

/* force definitions */
#define PANTEA_FORCE_REFERENCE 0 /* -dE_i/dr_i in the central role only == reference force.py:16-43 */
#define PANTEA_FORCE_FULL 1      /* -dE/dr_i of the total energy: every centre's contributions are also scattered to
                                    its neighbours (Newton's third law holds).  Not a reference mode: the reference
                                    differentiates the central copy of the positions only. */

typedef struct pantea_symfunc_desc {
    int32_t kind;        /* PANTEA_G1 / G2 / G3 / G9 */
    int32_t cutoff_type; /* PANTEA_CUT_* */
    int32_t type_j;      /* neighbour atom type (1-based) */
    int32_t type_k;      /* second neighbour type for G3/G9, 0 otherwise */
    double r_cutoff;
    double eta;
    double r_shift;      /* used by G2 only (G3/G9 ignore it, reference angular.py:58-65,100-107) */
    double lambda0;
    double zeta;
} pantea_symfunc_desc;

typedef struct pantea_element_desc {
    int32_t n_symfunc;                  /* radial functions first, then angular: output column order */
    const pantea_symfunc_desc* symfunc;
    /* scaler as x' = offset + slope * (x - shift) per feature; all NULL = identity
       (reference scaler.py:206-246 expressed as an affine map) */
    const double* scale_shift;
    const double* scale_slope;
    const double* scale_offset;
    /* network: n_layers dense layers, layer_sizes[n_layers+1] with layer_sizes[0] == n_symfunc,
       activations[n_layers], weights packed per layer as kernel [in,out] row-major then bias [out]
       (reference model.py:40-58; tests/test_nn.py:97-138 for the layout).  n_layers == 0: descriptor only. */
    int32_t n_layers;
    const int32_t* layer_sizes;
    const int32_t* activations;
    const double* weights;
} pantea_element_desc;

typedef struct pantea_potential_desc {
    int32_t n_elements;                  /* element e describes atom type e+1 */
    const pantea_element_desc* elements;
} pantea_potential_desc;

typedef struct pantea_potential pantea_potential;
typedef struct pantea_workspace pantea_workspace;

const char* pantea_last_error(void);
const char* pantea_version(void);
/* number of visible CUDA devices (0 when none / driver missing); never fails */
int pantea_device_count(void);

/* -- potential: replaces the static `atomic_potentials` pytree + params handed to the jitted kernels
      (reference potential.py:47-65, 140-317; energy.py:45-66).  Host descriptors are copied. */
int pantea_potential_create(const pantea_potential_desc* desc, pantea_potential** out);
int pantea_potential_destroy(pantea_potential* pot);
double pantea_potential_cutoff(const pantea_potential* pot); /* max r_cutoff (potential.py:374-382) */

/* -- workspace: library-owned scratch (cell list, neighbour rows, packed atom records) for up to
      `max_atoms` atoms and `max_neighbors` neighbours per atom.  `pot` may be NULL for a pure
      neighbour-search workspace. */
int pantea_workspace_create(const pantea_potential* pot, int64_t max_atoms, int32_t max_neighbors, int32_t dtype,
                            pantea_workspace** out);
int pantea_workspace_destroy(pantea_workspace* ws);

/* -- neighbour search: replaces `_calculate_cutoff_masks*` / `_calculate_distances*`
      (reference neighbor.py:74-115, distance.py:63-105, box.py:112-117).
      Neighbour <=> 0 < r <= r_cutoff with r = sqrt((dx^2+dy^2)+dz^2) of the single-shift minimum image
      d = r_i - r_j.  `box` = HOST double[3] lattice diagonal, or NULL for an open (non-periodic) structure.
      A periodic cell list is used when the box holds >= 3 cells of width >= r_cutoff per dimension,
      otherwise all pairs are scanned (exactly one image per pair, as the reference).
      Binds `positions`/`types` (snapshot) to the workspace for the calls below. */
int pantea_neighbor_build(pantea_workspace* ws, const void* positions /*[n,3]*/, const int32_t* types /*[n]*/,
                          int64_t n_atoms, const double* box, double r_cutoff, void* stream);
/* many independent structures in one launch (dataset preprocessing): atoms of structure s are
   [struct_ptr[s], struct_ptr[s+1]); boxes = DEVICE double [n_structs,3] or NULL. */
int pantea_neighbor_build_batch(pantea_workspace* ws, const void* positions, const int32_t* types, int64_t n_atoms,
                                const int32_t* struct_ptr /*[n_structs+1]*/, const double* boxes, int64_t n_structs,
                                double r_cutoff, void* stream);
/* restrict subsequent row building / energy evaluation to atoms [begin,end) (multi-GPU ownership) */
int pantea_workspace_set_owned_range(pantea_workspace* ws, int64_t begin, int64_t end);
/* synchronises `stream`; returns PANTEA_ECAPACITY if a row overflowed max_neighbors.  *max_count (HOST,
   may be NULL) receives the largest neighbour count seen. */
int pantea_neighbor_status(pantea_workspace* ws, int32_t* max_count, void* stream);
int pantea_neighbor_counts(pantea_workspace* ws, int32_t* counts /*[n]*/, void* stream);
/* CSR export with columns ascending within a row; row_ptr = exclusive prefix sum of the counts */
int pantea_neighbor_export(pantea_workspace* ws, const int64_t* row_ptr /*[n+1]*/, int32_t* col_idx, void* stream);
/* dense minimum-image distances (and optionally d = r_i - r_j) between two index subsets of the bound
   structure; NULL index = all atoms.  Replaces `calculate_distances` (reference distance.py:17-60). */
int pantea_distances(pantea_workspace* ws, const int32_t* idx_i, int64_t n_i, const int32_t* idx_j, int64_t n_j,
                     void* r /*[n_i,n_j]*/, void* d /*[n_i,n_j,3] or NULL*/, void* stream);

/* -- descriptor: replaces `_jitted_calculate_acsf_descriptor` and `_jitted_calculate_grad_acsf_descriptor`
      (reference acsf.py:163-228).  Applies the symmetry functions of element `element` (0-based slot) to the
      centres `centres[n_centres]` (atom indices of the bound structure, any type; NULL = all atoms).
      G [n_centres, n_symfunc]; dG [n_centres, n_symfunc, 3] = dG_i/dr_i in the central role, or NULL. */
int pantea_acsf_compute(pantea_workspace* ws, int32_t element, const int32_t* centres, int64_t n_centres, void* G,
                        void* dG, void* stream);

/* -- energy and forces: replaces `_jitted_compute_energy` and `_jitted_grad_compute_energy`
      (reference energy.py:45-66, force.py:16-43, potential.py:67-102).  One fused launch per call:
      symmetry functions + central gradients -> scaler -> per-element network forward/backward ->
      F_i = -sum_s dE_i/dG_is dG_is/dr_i.  e_atom [n] / forces [n,3] / e_total [1] may each be NULL.
      Only the owned range is written; e_total sums the owned atoms (deterministic order).
      force_mode PANTEA_FORCE_FULL: symmetry-function values -> networks -> a second pass over the same pair lists that
      scatters dE_i/dG_is dG_is/d(r_i - r_j) to the centre AND to its neighbours (shared-memory accumulation per neighbour,
      then one global atomic per neighbour and component; summation order, hence the last bits, not reproducible).
      All n rows of `forces` are zeroed and written: rows outside the owned range receive the owned centres' contributions
      to them (a multi-GPU caller sums those back onto their owners: pantea_halo_unpack_add). */
int pantea_energy_forces(pantea_workspace* ws, void* e_atom, void* forces, void* e_total, int32_t force_mode,
                         void* stream);

/* Evaluations that qualify -- PANTEA_F64, gradients requested, cell-list rows of a box that needs no minimum image on
   r_jk (every box length >= 4 r_cutoff), at least 64 atoms per SM, a potential with one tanhu cutoff per element whose
   angular groups are single G3 members with integer zeta -- run on specialised kernels (csrc/acsf2.cu: per-neighbour
   radial weights, r_jk from the unit-vector dot product).  Same results within ~1e-13 relative.  enable = 0 keeps every
   call on the generic kernels (diagnostics / parity tests); returns the previous setting.  Process-wide. */
int pantea_set_fast_path(int32_t enable);
/* Mixed precision of a PANTEA_F64 workspace (reference: `default_dtype.FLOATX = float32`, types.py:13-30, selects single
   precision for everything; here the state stays double so that large boxes lose nothing): bits = 32 evaluates the
   symmetry functions and their gradients in single precision -- the difference vectors r_i - r_j are still formed in
   double -- on the specialised kernels (same eligibility as pantea_set_fast_path; otherwise the setting has no effect);
   scaler, networks, forces and the integrator stay double.  Results agree with the double evaluation to ~1e-6
   relative (tolerance of the mode: 1e-5).  bits = 64 (default) restores the double evaluation. */
int pantea_workspace_set_compute_precision(pantea_workspace* ws, int32_t bits);
/* Gaussian screening of the specialised kernels: within one angular group of one centre, a triplet whose Gaussian weight
   exp(-eta (r_ij^2 + r_ik^2 + r_jk^2)) is below exp(-threshold) times the weight of the pair formed by the centre's two
   nearest neighbours of the group's types is not evaluated (it is dropped when the pair lists are built).  Active only
   when those two neighbours lie within r_cutoff of each other (the reference pair is then a live triplet) and no Verlet
   skin is set.  With the default threshold 40 (4e-18) the dropped terms of a centre sum to < 1e-14 of the symmetry
   function; 0 disables.  Returns the previous threshold.  Process-wide. */
double pantea_set_gauss_screen(double threshold);

/* -- molecular dynamics pieces: replace `_get_verlet_new_positions/_velocities`, `_wrap_into_box`,
      `_get_kinetic_energy`, `_get_rescaled_velocities` (reference molecular_dynamics.py:16-30, box.py:123-126,
      system.py:20-29, thermostat.py:12-22).  No mass enters the integrator (as the reference).
      All act on atoms [begin,end). */
int pantea_md_update_positions(void* positions, const void* velocities, const void* forces, int64_t begin,
                               int64_t end, const double* box, double dt, int32_t dtype, void* stream);
int pantea_md_update_velocities(void* velocities, void* forces /*in: F(t), out: F(t+dt)*/, const void* new_forces,
                                int64_t begin, int64_t end, double dt, int32_t dtype, void* stream);
/* Extension ("mass_scaled physical MD"): the same two updates with the acceleration F / m in place of
   F; masses [n] in the workspace dtype, NULL = the reference update.  With PANTEA_FORCE_FULL forces this is the usual
   energy-conserving velocity Verlet. */
int pantea_md_update_positions_mass(void* positions, const void* velocities, const void* forces, const void* masses,
                                    int64_t begin, int64_t end, const double* box, double dt, int32_t dtype, void* stream);
int pantea_md_update_velocities_mass(void* velocities, void* forces, const void* new_forces, const void* masses,
                                     int64_t begin, int64_t end, double dt, int32_t dtype, void* stream);
/* ke_out: DEVICE double[1] = 0.5 * sum m v^2 over [begin,end) (fixed-order reduction) */
int pantea_md_kinetic_energy(const void* velocities, const void* masses, int64_t begin, int64_t end, double* ke_out,
                             int32_t dtype, void* stream);
/* Berendsen: v *= 1/sqrt(1 + dt/tau (T/T0 - 1)), T = 2 KE / (3 n_total kB), KE read from DEVICE ke[0] */
int pantea_md_rescale_velocities(void* velocities, int64_t begin, int64_t end, const double* ke, int64_t n_total,
                                 double dt, double tau, double t_target, double kb, int32_t dtype, void* stream);

typedef struct pantea_md_params {
    double dt;
    double t_target;  /* Berendsen target temperature [K]; ignored when tau <= 0 */
    double tau;       /* Berendsen time constant; <= 0 disables the thermostat */
    double kb;        /* Boltzmann constant in the caller's units */
    int32_t record;   /* != 0: write (E_pot, E_kin) of every step into `scalars` */
    int32_t use_graph;/* != 0: replay the step as a CUDA graph */
    int32_t mass_scaled; /* != 0: extension -- accelerations F/m in both half-steps (the reference integrator has none) */
    int32_t force_mode;  /* PANTEA_FORCE_REFERENCE (0, the reference) or PANTEA_FORCE_FULL */
} pantea_md_params;

/* Runs n_steps velocity-Verlet steps entirely on the device, no host synchronisation inside:
   replaces the `simulate` -> `MDSimulator.simulate_one_step` loop (reference simulate.py:77-83,
   molecular_dynamics.py:57-77).  forces must hold F(positions) on entry.  scalars: DEVICE double
   [n_steps, 2] (E_pot, E_kin after each step) when params->record, else may be NULL. */
int pantea_md_run(pantea_workspace* ws, void* positions, void* velocities, void* forces, const void* masses,
                  const int32_t* types, int64_t n_atoms, const double* box, int64_t n_steps,
                  const pantea_md_params* params, double* scalars, void* stream);

/* -- Lennard-Jones potential on the bound neighbour rows: replaces `_jitted_compute_total_energy` and `_compute_forces`
      (reference simulation/lennard_jones.py:73-123).  forces = +dE/dr_i, as the reference returns (sic).
      e_atom [n] (half of each pair energy per atom) / forces [n,3] / e_total [1] may each be NULL. */
int pantea_lj_energy_forces(pantea_workspace* ws, double sigma, double epsilon, void* e_atom, void* forces,
                            void* e_total, void* stream);

/* -- measurement helpers (bench.py) ------------------------------------------------------------ */
/* number of kernel launches issued through this library since load */
int64_t pantea_launch_count(void);
/* FMA-pipe peak microbenchmark (8 independent chains per thread); *flops = flops of the launch (HOST). */
int pantea_bench_fma(int32_t dtype, int32_t iters, int32_t blocks, int32_t threads, void* scratch, double* flops,
                     void* stream);
/* overwrite `bytes` (> L2 capacity) of DEVICE scratch so that the following kernel sees a cold L2 */
int pantea_l2_flush(void* scratch, int64_t bytes, void* stream);
/* enable != 0: bracket every later launch of the specialised evaluation kernel (the dominant kernel of the step) with
   CUDA events on its own stream; *ms (HOST, may be NULL) = duration of the most recent bracketed launch.  Eager
   launches only (not inside graph capture). */
int pantea_eval_timing(int32_t enable, float* ms);
/* optional work counters, DEVICE uint64[4] or NULL: [0] neighbour pairs, [1] radial-SF evaluations,
   [2] triplet-SF evaluations (generic kernels: live triplets of the reference algorithm), [3] pair-list entries the
   specialised evaluation walks (after Gaussian screening, without padding) -- accumulated by every later descriptor /
   energy launch of `ws` */
int pantea_workspace_set_counters(pantea_workspace* ws, void* counters);

/* Verlet-skin reuse of the neighbour rows and pair lists (the reference notes the missing
   feature in atoms/neighbor.py:34-36).  With skin > 0 and a periodic cell-list build, pantea_neighbor_build gathers
   rows with radius r_cutoff + skin and later calls with the same atoms / box / cutoff rebuild them only when some atom
   has moved more than skin / 2 (minimum image) since the last rebuild; the decision is taken on the device, so the call
   sequence is static (CUDA-graph safe).  Neighbours beyond a symmetry function's cutoff contribute exactly zero, so
   energies and forces agree with skin = 0 up to summation order.  The exact-set queries (pantea_neighbor_counts /
   _export) and pantea_lj_energy_forces refuse rows built with a skin (PANTEA_EINVAL).  skin = 0 (default) disables. */
int pantea_workspace_set_skin(pantea_workspace* ws, double skin);
/* builds[0] = pantea_neighbor_build calls that ran with a skin, builds[1] = how many of them rebuilt the rows
   (synchronises `stream`) */
int pantea_neighbor_rebuilds(pantea_workspace* ws, int64_t* builds, void* stream);

/* Per-feature statistics of a descriptor batch: data is a row-major DEVICE matrix [n_rows, n_cols] (row stride `ld`
   elements, PANTEA_F32 / PANTEA_F64); stats is DEVICE double[4 * n_cols] = mean | population sigma (two-pass) | min |
   max.  Replaces the jnp.mean / jnp.std / jnp.min / jnp.max reductions of DescriptorScaler.fit / partial_fit
   (descriptors/scaler.py:250-283), i.e. the work of trainer.fit_scaler (potentials/nnp/trainer.py:68-88).  Fixed-shape
   two-stage reductions: bitwise reproducible. */
int pantea_scaler_stats(const void* data, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t dtype, double* stats,
                        void* stream);

/* -- halo exchange of the brick-decomposed MD; the transport itself is NCCL send/recv issued by the
      caller (torch.distributed all_to_all_single), these are the device-side pack / unpack steps.
   pantea_halo_pack: send_buf[i] = positions[send_idx[i]] for the fixed send list of the current ghost shell, and -- when
      pos_ref is given -- raise the sticky DEVICE flag *violated if an owned atom [0,n_own) moved more than `limit`
      (= skin / 2; nearest image in the HOST box, NULL = open) away from pos_ref, the positions the lists were built on.
   pantea_halo_unpack_add: out[a] += sum of recv_buf[e] over the entries e of the send list that refer to owned atom a
      (reverse halo of PANTEA_FORCE_FULL); order [n_send] = send-list entries sorted by atom (stable), first [n_own+1] =
      each atom's range in `order`.  One thread per atom, fixed order: reproducible. */
int pantea_halo_pack(const void* positions /*[n,3]*/, const int64_t* send_idx, int64_t n_send, void* send_buf /*[n_send,3]*/,
                     const void* pos_ref /*[n_own,3] or NULL*/, int64_t n_own, const double* box, double limit,
                     int32_t* violated, int32_t dtype, void* stream);
int pantea_halo_unpack_add(void* out /*[n_own,3]*/, const void* recv_buf /*[n_send,3]*/, const int64_t* order,
                           const int64_t* first, int64_t n_own, int32_t dtype, void* stream);

/* -- brick-decomposed MD over NVLink peer memory ("3-D spatial bricks, ghost shell width = rc, forward
      halo of ghost positions, atom migration"; the reference is single-process: atoms/structure.py:45-49 only mentions
      the idea).  One process per GPU.  Atoms keep their global index on every rank; a rank owns the atoms inside its
      brick and holds ghost copies of those within r_cutoff of it.  Per step: ONE kernel integrates the owned atoms and
      stores their new positions straight into the mailbox rows of every rank that needs them (st.global on CUDA-IPC
      mapped peer pointers = NVLink writes; migrating atoms take their velocity and force rows along) and publishes the
      step number; ONE kernel waits for all peers' step numbers and unpacks; then the single-GPU pipeline runs on the
      present atoms (global cell grid: per-atom results do not depend on the number of ranks); then the velocity
      update.  Captured once as a CUDA graph; no collective library call and no host synchronisation inside.
      dims[3] = bricks per axis (product = world, rank = (ix*py+iy)*pz+iz), cuts_d = the dims[d]-1 interior boundaries
      (HOST, ascending; NULL when dims[d] == 1), own_cap = upper bound of the atoms one rank may own (launch grids).
      Integrator = the reference's (no mass, no thermostat: NVE). */
typedef struct pantea_mgpu pantea_mgpu;
int pantea_mgpu_create(pantea_workspace* ws, int32_t rank, int32_t world, int64_t n_atoms, const double* box,
                       const int32_t* dims, const double* cuts_x, const double* cuts_y, const double* cuts_z,
                       double r_cutoff, double dt, int64_t own_cap, pantea_mgpu** out);
/* peer mapping: every rank exports pantea_mgpu_handle_bytes() bytes (a cudaIpcMemHandle_t), the caller gathers them in
   rank order (HOST) and hands the concatenation to pantea_mgpu_connect (world == 1: connected at creation) */
int64_t pantea_mgpu_handle_bytes(void);
int pantea_mgpu_export_handle(pantea_mgpu* mg, void* handle_out);
int pantea_mgpu_connect(pantea_mgpu* mg, const void* all_handles);
/* (re)start from full-length replicated DEVICE arrays positions / velocities [n,3] (same on every rank; types [n] is
   borrowed for the lifetime of the runs) and compute the owned atoms' forces.  The caller puts a barrier (all ranks
   idle) before and after this call. */
int pantea_mgpu_set_state(pantea_mgpu* mg, const void* positions, const void* velocities, const int32_t* types, void* stream);
int pantea_mgpu_run(pantea_mgpu* mg, int64_t n_steps, int32_t use_graph, void* stream);
/* copies of the rank's full-length arrays (any may be NULL): positions are valid where roles >= 1, velocities / forces
   where roles == 2 (0 absent, 1 ghost, 2 owned).  *status (HOST, may be NULL; synchronises): 0 ok, 1 + r = peer r never
   published a step within the spin limit, 1000 = own_cap exceeded. */
int pantea_mgpu_read(pantea_mgpu* mg, void* positions, void* velocities, void* forces, uint8_t* roles, int32_t* status,
                     void* stream);
int pantea_mgpu_destroy(pantea_mgpu* mg);

#ifdef __cplusplus
}
#endif
#endif /* PANTEA_B200_H */
