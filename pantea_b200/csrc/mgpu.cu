// Brick-decomposed molecular dynamics over NVLink peer memory (north star "the box is spatially
// domain-decomposed across the GPUs of one box, with ghost-atom halo exchange").  One process per GPU; the reference is
// single-process and has no counterpart (pantea/atoms/structure.py:45-49 is a docstring).
//
// The periodic box is cut into px x py x pz bricks.  A rank OWNS the atoms whose wrapped position lies in its brick and
// needs, as ghosts, the foreign atoms within r_cutoff of the brick.  Atoms keep their global index on every rank: each
// rank holds full-length arrays of which only the rows of its owned (velocities, forces) or owned + ghost (positions)
// atoms are meaningful, and a per-atom role (0 absent, 1 ghost, 2 owned) that the neighbour kernels honour
// (pantea_workspace.role).  No lists, no packing, no counts travel:
//
//   1. integrate + push (ONE kernel).  Every owner advances its atoms (the reference's mass-less velocity Verlet,
//      molecular_dynamics.py:16-30, wrap of box.py:123-126) and stores the new position, stamped with the step number,
//      straight into the mailbox row of that atom on every rank whose brick + r_cutoff shell contains it -- plain
//      st.global on peer pointers mapped with CUDA IPC, i.e. NVLink / NVSwitch writes issued by the integration kernel
//      itself; only halo atoms cross a link.  An atom that leaves the brick also takes its velocity and force rows to
//      the new owner.
//   2. wait + unpack (ONE kernel).  Its first thread publishes the step number in every peer's flag row (system-scope
//      release: the integration kernel has completed, its peer stores are performed).  Blocks then wait (bounded spin, system-scope acquire) until every peer has published
//      this step, then turn the stamped mailbox rows into the dense position array and the roles of this step; rows
//      with an older stamp are atoms that are not in this rank's shell any more.
//   3. the unchanged single-GPU pipeline on the present atoms: binning into the GLOBAL cell grid (so rows, and hence
//      every bit of a per-atom result, do not depend on the number of ranks), neighbour rows / pair lists / symmetry
//      functions / networks for the owned atoms only.  The reference force is the central-role derivative
//      (force.py:16-43): no reverse communication.
//   4. velocity update of the owned atoms.
//
// The mailbox is double-buffered by step parity: a rank can only run one step ahead of its slowest peer (it waits for
// the peer's flag of step s before it reads, and the peer published that flag after it finished reading step s - 1).
// The whole step is captured once as a CUDA graph and replayed: no host synchronisation, no collective library call.
#include <cstring>

#include "internal.cuh"
#include "math.cuh"

namespace pantea {

constexpr int kMaxRanks = 16;

template <typename T> struct MailRec;
template <> struct __align__(16) MailRec<double> { double x, y, z; long long stamp; };
template <> struct __align__(16) MailRec<float> { float x, y, z; int stamp; };

struct BrickTable {
    int rank, world;
    int dims[3];
    double box[3];
    double edges[3][kMaxRanks + 1];  // brick boundaries per axis: edges[d][0] = 0 ... edges[d][dims[d]] = box[d]
    double reach;                    // ghost shell width (r_cutoff, slightly inclusive)
};

struct PeerTable {
    void* mail[kMaxRanks];              // MailRec<T>[2][n]
    void* vel[kMaxRanks];               // T[n][3]
    void* frc[kMaxRanks];               // T[n][3]
    unsigned long long* flags[kMaxRanks];  // [world] step numbers published by the peers
};

__device__ __forceinline__ int brick_index(const BrickTable& b, int d, double x) {
    int k = 0;
    for (int c = 1; c < b.dims[d]; ++c) k += x >= b.edges[d][c] ? 1 : 0;  // a coordinate on a boundary belongs to the upper brick
    return k;
}
__device__ __forceinline__ int brick_owner(const BrickTable& b, double x, double y, double z) {
    return (brick_index(b, 0, x) * b.dims[1] + brick_index(b, 1, y)) * b.dims[2] + brick_index(b, 2, z);
}
// periodic distance of coordinate x to the interval [lo, hi) along an axis of length L (0 inside)
__device__ __forceinline__ double axis_gap(double x, double lo, double hi, double L) {
    if (x >= lo && x < hi) return 0.0;
    double a = lo - x; if (a < 0.0) a += L;  // forward distance to the lower edge (x is wrapped: |lo - x| < L)
    double c = x - hi; if (c < 0.0) c += L;  // backward distance to the upper edge
    return a < c ? a : c;
}
__device__ __forceinline__ bool in_shell(const BrickTable& b, int r, double x, double y, double z) {
    const int pz = b.dims[2], py = b.dims[1];
    const int c[3] = {r / (py * pz), (r / pz) % py, r % pz};
    const double p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (b.dims[d] == 1) continue;  // the brick spans the whole periodic axis
        if (axis_gap(p[d], b.edges[d][c[d]], b.edges[d][c[d] + 1], b.box[d]) > b.reach) return false;
    }
    return true;
}

template <typename T> __device__ __forceinline__ T t_fmod2(T a, T b);
template <> __device__ __forceinline__ double t_fmod2<double>(double a, double b) { return fmod(a, b); }
template <> __device__ __forceinline__ float t_fmod2<float>(float a, float b) { return fmodf(a, b); }
template <typename T>
__device__ __forceinline__ T wrap1(T x, T box) {  // floored remainder, as md.cu / box.py:123-126
    if (x >= (T)0 && x < box) return x;  // already inside: the remainder is x itself (fmod is the slow path)
    T m = t_fmod2<T>(x, box);
    if (m != (T)0 && ((m < (T)0) != (box < (T)0))) m = add_rn(m, box);
    return m;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// 1. integrate + push
template <typename T>
__global__ void mgpu_integrate_push_kernel(int n, const uint8_t* __restrict__ role, const T* __restrict__ pos,
                                           const BrickTable* __restrict__ bt_p, const PeerTable* __restrict__ pt_p,
                                           const unsigned long long* __restrict__ epoch_p, T dt) {
    const BrickTable& bt = *bt_p;
    const PeerTable& pt = *pt_p;
    const unsigned long long next = *epoch_p + 1;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    // brick bounds of every rank, once per block
    __shared__ double s_lo[kMaxRanks][3], s_hi[kMaxRanks][3];
    if (threadIdx.x < bt.world) {
        const int r = threadIdx.x, pz = bt.dims[2], py = bt.dims[1];
        const int c[3] = {r / (py * pz), (r / pz) % py, r % pz};
        for (int d = 0; d < 3; ++d) { s_lo[r][d] = bt.edges[d][c[d]]; s_hi[r][d] = bt.edges[d][c[d] + 1]; }
    }
    __syncthreads();
    if (id < n && role[id] == 2) {
        const T* vel = (const T*)pt.vel[bt.rank];
        const T* frc = (const T*)pt.frc[bt.rank];
        T x[3], v[3], f[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            v[d] = vel[3 * id + d]; f[d] = frc[3 * id + d];
            // x + v*dt + 0.5*F*dt*dt evaluated left to right without contraction (molecular_dynamics.py:20), then wrapped
            x[d] = add_rn(add_rn(pos[3 * id + d], mul_rn(v[d], dt)), mul_rn(mul_rn(mul_rn((T)0.5, f[d]), dt), dt));
            x[d] = wrap1<T>(x[d], (T)bt.box[d]);
        }
        MailRec<T> m;
        m.x = x[0]; m.y = x[1]; m.z = x[2]; m.stamp = (decltype(m.stamp))next;
        const int owner = brick_owner(bt, (double)x[0], (double)x[1], (double)x[2]);
        for (int r = 0; r < bt.world; ++r) {
            bool inside = true;
#pragma unroll
            for (int d = 0; d < 3; ++d)
                if (bt.dims[d] > 1 && axis_gap((double)x[d], s_lo[r][d], s_hi[r][d], bt.box[d]) > bt.reach) inside = false;
            if (!inside) continue;
            ((MailRec<T>*)pt.mail[r])[(size_t)(next & 1) * n + id] = m;  // NVLink store when r is a peer
            if (r == owner && r != bt.rank) {  // the atom changes its owner: velocity and force rows travel with it
                T* pv = (T*)pt.vel[r];
                T* pf = (T*)pt.frc[r];
#pragma unroll
                for (int d = 0; d < 3; ++d) { pv[3 * id + d] = v[d]; pf[3 * id + d] = f[d]; }
            }
        }
    }
}

// 2. wait + unpack
template <typename T>
__global__ void mgpu_wait_unpack_kernel(int n, uint8_t* __restrict__ role, T* __restrict__ pos,
                                        const BrickTable* __restrict__ bt_p, const PeerTable* __restrict__ pt_p,
                                        const unsigned long long* __restrict__ epoch_p, int advance,
                                        long long spin_limit, int* __restrict__ err, unsigned int* __restrict__ done) {
    const BrickTable& bt = *bt_p;
    const PeerTable& pt = *pt_p;
    const unsigned long long want = *epoch_p + (advance ? 1 : 0);
    // publish: the integrate + push kernel before this one has completed (stream order), so its stores -- also those to
    // peer memory -- are performed; one thread makes that a system-scope release of this rank's step number
    // (the first block to RUN does it, whichever it is: the blocks that wait below must not be able to keep it off the SMs)
    if (threadIdx.x == 0 && advance) {
        const unsigned int ticket = atomicAdd(done, 1u);
        if (ticket == 0) {
            __threadfence_system();
            for (int r = 0; r < bt.world; ++r) st_release_sys(pt.flags[r] + bt.rank, want);
        }
        if (ticket == gridDim.x - 1) *done = 0;
        const unsigned long long* flags = pt.flags[bt.rank];
        const long long t0 = clock64();
        for (int r = 0; r < bt.world; ++r) {
            if (r == bt.rank) continue;  // own rows are local and complete (stream order)
            while (ld_acquire_sys(flags + r) < want) {
                if (clock64() - t0 > spin_limit) { atomicExch(err, 1 + r); break; }  // a peer never arrived: give up, report
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const MailRec<T> m = ((const MailRec<T>*)pt.mail[bt.rank])[(size_t)(want & 1) * n + id];
    uint8_t r = 0;
    if ((unsigned long long)m.stamp == want) {
        pos[3 * id] = m.x; pos[3 * id + 1] = m.y; pos[3 * id + 2] = m.z;
        r = brick_owner(bt, (double)m.x, (double)m.y, (double)m.z) == bt.rank ? 2 : 1;
    }
    role[id] = r;
}

// 4. velocity update of the owned atoms; advances the step number
template <typename T>
__global__ void mgpu_velocities_kernel(int n, const uint8_t* __restrict__ role, const BrickTable* __restrict__ bt_p,
                                       const PeerTable* __restrict__ pt_p, const T* __restrict__ frc_new, T dt,
                                       unsigned long long* __restrict__ epoch_p, const int32_t* __restrict__ n_owned,
                                       int own_cap, int* __restrict__ err) {
    const BrickTable& bt = *bt_p;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id < n && role[id] == 2) {
        T* vel = (T*)pt_p->vel[bt.rank];
        T* frc = (T*)pt_p->frc[bt.rank];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const T fn = frc_new[3 * id + d];
            vel[3 * id + d] = add_rn(vel[3 * id + d], mul_rn(mul_rn((T)0.5, add_rn(frc[3 * id + d], fn)), dt));  // molecular_dynamics.py:30
            frc[3 * id + d] = fn;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *epoch_p += 1;
        if (n_owned && *n_owned > own_cap) atomicExch(err, 1000);  // more owned atoms than the launch grids cover: results invalid
    }
}

// initial / reset state: every atom goes into the own mailbox with the current stamp
template <typename T>
__global__ void mgpu_fill_mail_kernel(int n, const T* __restrict__ pos_in, const T* __restrict__ vel_in,
                                      const BrickTable* __restrict__ bt_p, const PeerTable* __restrict__ pt_p,
                                      const unsigned long long* __restrict__ epoch_p) {
    const BrickTable& bt = *bt_p;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const unsigned long long e = *epoch_p;
    MailRec<T> m;
    m.x = wrap1<T>(pos_in[3 * id], (T)bt.box[0]); m.y = wrap1<T>(pos_in[3 * id + 1], (T)bt.box[1]);
    m.z = wrap1<T>(pos_in[3 * id + 2], (T)bt.box[2]); m.stamp = (decltype(m.stamp))e;
    MailRec<T>* mail = (MailRec<T>*)pt_p->mail[bt.rank];
    mail[(size_t)(e & 1) * n + id] = m;
    m.stamp = (decltype(m.stamp))(-1);
    mail[(size_t)((e + 1) & 1) * n + id] = m;  // the other half holds nothing valid
    T* vel = (T*)pt_p->vel[bt.rank];
#pragma unroll
    for (int d = 0; d < 3; ++d) vel[3 * id + d] = vel_in[3 * id + d];
}

__global__ void mgpu_copy_role_forces_kernel(int n, const uint8_t* __restrict__ role, const void* __restrict__ src,
                                             void* __restrict__ dst, int esz) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n || role[id] != 2) return;
    for (int b = 0; b < 3 * esz; ++b) ((unsigned char*)dst)[(size_t)3 * esz * id + b] = ((const unsigned char*)src)[(size_t)3 * esz * id + b];
}

}  // namespace pantea

using namespace pantea;

struct pantea_mgpu {
    pantea_workspace* ws = nullptr;
    int rank = 0, world = 1, dtype = PANTEA_F64;
    int64_t n = 0;
    double rc = 0, dt = 0;
    BrickTable bt_host;
    PeerTable pt_host;
    void* block = nullptr;  // exported allocation: mailbox | velocities | forces | flags
    size_t off_vel = 0, off_frc = 0, off_flags = 0, block_bytes = 0;
    void* peer_block[kMaxRanks] = {nullptr};
    bool connected = false;
    BrickTable* bt_dev = nullptr;
    PeerTable* pt_dev = nullptr;
    void* pos = nullptr;       // [n,3] dense positions of the present atoms
    void* frc_new = nullptr;   // [n,3]
    uint8_t* role = nullptr;   // [n]
    const int32_t* types = nullptr;  // borrowed
    unsigned long long* epoch = nullptr;
    unsigned int* done = nullptr;
    int* err = nullptr;
    cudaGraphExec_t graph = nullptr;
    cudaStream_t capture_stream = nullptr;
    int64_t graph_epoch = -1;
    int graph_nodes = 0;
};

namespace {

size_t esize(int dtype) { return dtype == PANTEA_F64 ? 8 : 4; }

template <typename T>
int mgpu_step_typed(pantea_mgpu* mg, cudaStream_t st) {
    const int n = (int)mg->n, threads = 256, blocks = (n + threads - 1) / threads;
    mgpu_integrate_push_kernel<T><<<(n + 511) / 512, 512, 0, st>>>(n, mg->role, (const T*)mg->pos, mg->bt_dev, mg->pt_dev, mg->epoch,
                                                              (T)mg->dt);
    PANTEA_LAUNCH_CHECK();
    mgpu_wait_unpack_kernel<T><<<blocks, threads, 0, st>>>(n, mg->role, (T*)mg->pos, mg->bt_dev, mg->pt_dev, mg->epoch, 1,
                                                           (long long)2e10, mg->err, mg->done);
    PANTEA_LAUNCH_CHECK();
    int rc = neighbor_build_impl(mg->ws, mg->pos, mg->types, mg->n, mg->bt_host.box, nullptr, nullptr, 1, mg->rc, st);
    if (rc) return rc;
    rc = atom_kernel_launch(mg->ws, -1, nullptr, 0, nullptr, nullptr, nullptr, mg->frc_new, st, PANTEA_FORCE_REFERENCE);
    if (rc) return rc;
    const pantea_workspace* ws = mg->ws;
    const int32_t* n_owned = ws->role ? ws->cell_own + (int64_t)ws->ncell[0] * ws->ncell[1] * ws->ncell[2] : nullptr;
    mgpu_velocities_kernel<T><<<blocks, threads, 0, st>>>(n, mg->role, mg->bt_dev, mg->pt_dev, (const T*)mg->frc_new, (T)mg->dt,
                                                          mg->epoch, n_owned, (int)ws->own_cap, mg->err);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int mgpu_step(pantea_mgpu* mg, cudaStream_t st) {
    return mg->dtype == PANTEA_F64 ? mgpu_step_typed<double>(mg, st) : mgpu_step_typed<float>(mg, st);
}

}  // namespace

extern "C" {

int pantea_mgpu_create(pantea_workspace* ws, int32_t rank, int32_t world, int64_t n_atoms, const double* box,
                       const int32_t* dims, const double* cuts_x, const double* cuts_y, const double* cuts_z,
                       double r_cutoff, double dt, int64_t own_cap, pantea_mgpu** out) {
    if (!ws || !ws->pot || !box || !dims || !out) return fail(PANTEA_EINVAL, "pantea_mgpu_create: NULL argument");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(PANTEA_EINVAL, "pantea_mgpu_create: bad rank / world");
    if (dims[0] * dims[1] * dims[2] != world) return fail(PANTEA_EINVAL, "pantea_mgpu_create: brick grid does not match the number of ranks");
    if (n_atoms < 1 || n_atoms > ws->max_atoms) return fail(PANTEA_EINVAL, "pantea_mgpu_create: n_atoms exceeds the workspace capacity");
    auto* mg = new pantea_mgpu();
    mg->ws = ws; mg->rank = rank; mg->world = world; mg->dtype = ws->dtype; mg->n = n_atoms; mg->rc = r_cutoff; mg->dt = dt;
    BrickTable& bt = mg->bt_host;
    std::memset(&bt, 0, sizeof(bt));
    bt.rank = rank; bt.world = world;
    const double* cuts[3] = {cuts_x, cuts_y, cuts_z};
    for (int d = 0; d < 3; ++d) {
        bt.dims[d] = dims[d]; bt.box[d] = box[d];
        bt.edges[d][0] = 0.0; bt.edges[d][dims[d]] = box[d];
        for (int c = 1; c < dims[d]; ++c) {
            if (!cuts[d]) { delete mg; return fail(PANTEA_EINVAL, "pantea_mgpu_create: missing brick boundaries"); }
            bt.edges[d][c] = cuts[d][c - 1];
        }
    }
    bt.reach = r_cutoff * (1.0 + 1e-9) + 1e-9;
    const size_t es = esize(mg->dtype), n = (size_t)n_atoms;
    const size_t mail_bytes = 2 * n * (mg->dtype == PANTEA_F64 ? sizeof(MailRec<double>) : sizeof(MailRec<float>));
    mg->off_vel = (mail_bytes + 255) & ~size_t(255);
    mg->off_frc = (mg->off_vel + 3 * n * es + 255) & ~size_t(255);
    mg->off_flags = (mg->off_frc + 3 * n * es + 255) & ~size_t(255);
    mg->block_bytes = mg->off_flags + 256;
    cudaError_t err = cudaMalloc(&mg->block, mg->block_bytes);
    if (err == cudaSuccess) err = cudaMemset(mg->block, 0, mg->block_bytes);
    auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(p, bytes); if (err == cudaSuccess) err = cudaMemset(*p, 0, bytes); };
    alloc(&mg->pos, 3 * n * es); alloc(&mg->frc_new, 3 * n * es); alloc((void**)&mg->role, n);
    alloc((void**)&mg->epoch, 8); alloc((void**)&mg->done, 4); alloc((void**)&mg->err, 4);
    alloc((void**)&mg->bt_dev, sizeof(BrickTable)); alloc((void**)&mg->pt_dev, sizeof(PeerTable));
    if (err != cudaSuccess) {
        pantea_mgpu_destroy(mg);
        return fail(err == cudaErrorMemoryAllocation ? PANTEA_ENOMEM : PANTEA_ECUDA, std::string("pantea_mgpu_create: ") + cudaGetErrorString(err));
    }
    mg->peer_block[rank] = mg->block;
    PANTEA_CUDA_TRY(cudaMemcpy(mg->bt_dev, &bt, sizeof(bt), cudaMemcpyHostToDevice));
    if (world == 1) {
        int rc = pantea_mgpu_connect(mg, nullptr);
        if (rc) { pantea_mgpu_destroy(mg); return rc; }
    }
    // rows and evaluation run over the owned atoms only; their number is known on the device, `own_cap` bounds it for
    // the launch grids (exceeding it raises the error status)
    ws->role = world > 1 ? mg->role : nullptr;  // one rank owns every atom: the plain single-GPU pipeline
    ws->own_cap = own_cap > 0 && own_cap < n_atoms ? own_cap : n_atoms;
    ws->own_begin = 0; ws->own_end = -1;
    ++ws->arg_epoch;
    *out = mg;
    return PANTEA_OK;
}

int64_t pantea_mgpu_handle_bytes(void) { return (int64_t)sizeof(cudaIpcMemHandle_t); }

int pantea_mgpu_export_handle(pantea_mgpu* mg, void* handle_out) {
    if (!mg || !handle_out) return fail(PANTEA_EINVAL, "pantea_mgpu_export_handle: NULL argument");
    cudaIpcMemHandle_t h;
    PANTEA_CUDA_TRY(cudaIpcGetMemHandle(&h, mg->block));
    std::memcpy(handle_out, &h, sizeof(h));
    return PANTEA_OK;
}

int pantea_mgpu_connect(pantea_mgpu* mg, const void* all_handles) {
    if (!mg) return fail(PANTEA_EINVAL, "pantea_mgpu_connect: NULL argument");
    if (mg->world > 1 && !all_handles) return fail(PANTEA_EINVAL, "pantea_mgpu_connect: handles of all ranks are required");
    for (int r = 0; r < mg->world; ++r) {
        if (r == mg->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const unsigned char*)all_handles + (size_t)r * sizeof(h), sizeof(h));
        cudaError_t err = cudaIpcOpenMemHandle(&mg->peer_block[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess)
            return fail(PANTEA_ECUDA, std::string("pantea_mgpu_connect: cudaIpcOpenMemHandle (peer access over NVLink is required): ") + cudaGetErrorString(err));
    }
    PeerTable& pt = mg->pt_host;
    std::memset(&pt, 0, sizeof(pt));
    for (int r = 0; r < mg->world; ++r) {
        unsigned char* b = (unsigned char*)mg->peer_block[r];
        pt.mail[r] = b; pt.vel[r] = b + mg->off_vel; pt.frc[r] = b + mg->off_frc;
        pt.flags[r] = (unsigned long long*)(b + mg->off_flags);
    }
    PANTEA_CUDA_TRY(cudaMemcpy(mg->pt_dev, &pt, sizeof(pt), cudaMemcpyHostToDevice));
    mg->connected = true;
    return PANTEA_OK;
}

// (re)start from replicated full-length host-provided DEVICE arrays: positions, velocities [n,3]; types [n] is borrowed.
// Computes the forces of the owned atoms.  Every rank must call it with the same arrays, and the caller must put a
// barrier between this call and the first pantea_mgpu_run (a peer that is already running would write velocity rows
// of migrating atoms into arrays this call is still filling).
int pantea_mgpu_set_state(pantea_mgpu* mg, const void* positions, const void* velocities, const int32_t* types, void* stream) {
    if (!mg || !positions || !velocities || !types) return fail(PANTEA_EINVAL, "pantea_mgpu_set_state: NULL argument");
    if (!mg->connected) return fail(PANTEA_EINVAL, "pantea_mgpu_set_state: call pantea_mgpu_connect first");
    cudaStream_t st = (cudaStream_t)stream;
    mg->types = types;
    const int n = (int)mg->n, threads = 256, blocks = (n + threads - 1) / threads;
    int rc;
    if (mg->dtype == PANTEA_F64) {
        mgpu_fill_mail_kernel<double><<<blocks, threads, 0, st>>>(n, (const double*)positions, (const double*)velocities, mg->bt_dev, mg->pt_dev, mg->epoch);
        PANTEA_LAUNCH_CHECK();
        mgpu_wait_unpack_kernel<double><<<blocks, threads, 0, st>>>(n, mg->role, (double*)mg->pos, mg->bt_dev, mg->pt_dev, mg->epoch, 0, 0, mg->err, mg->done);
    } else {
        mgpu_fill_mail_kernel<float><<<blocks, threads, 0, st>>>(n, (const float*)positions, (const float*)velocities, mg->bt_dev, mg->pt_dev, mg->epoch);
        PANTEA_LAUNCH_CHECK();
        mgpu_wait_unpack_kernel<float><<<blocks, threads, 0, st>>>(n, mg->role, (float*)mg->pos, mg->bt_dev, mg->pt_dev, mg->epoch, 0, 0, mg->err, mg->done);
    }
    PANTEA_LAUNCH_CHECK();
    rc = neighbor_build_impl(mg->ws, mg->pos, mg->types, mg->n, mg->bt_host.box, nullptr, nullptr, 1, mg->rc, st);
    if (rc) return rc;
    rc = atom_kernel_launch(mg->ws, -1, nullptr, 0, nullptr, nullptr, nullptr, mg->frc_new, st, PANTEA_FORCE_REFERENCE);
    if (rc) return rc;
    mgpu_copy_role_forces_kernel<<<blocks, threads, 0, st>>>(n, mg->role, mg->frc_new, mg->pt_host.frc[mg->rank], (int)esize(mg->dtype));
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int pantea_mgpu_run(pantea_mgpu* mg, int64_t n_steps, int32_t use_graph, void* stream) {
    if (!mg) return fail(PANTEA_EINVAL, "pantea_mgpu_run: NULL argument");
    if (!mg->connected || !mg->types) return fail(PANTEA_EINVAL, "pantea_mgpu_run: connect and set the state first");
    if (n_steps <= 0) return PANTEA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t done = 0;
    if (!(use_graph && mg->graph && mg->graph_epoch == mg->ws->arg_epoch)) {
        // an eager step first: it performs every lazy allocation / attribute setting outside of graph capture
        int rc = mgpu_step(mg, st);
        if (rc) return rc;
        done = 1;
    }
    if (use_graph && n_steps - done >= 1) {
        if (!(mg->graph && mg->graph_epoch == mg->ws->arg_epoch)) {
            if (mg->graph) { cudaGraphExecDestroy(mg->graph); mg->graph = nullptr; }
            if (!mg->capture_stream) PANTEA_CUDA_TRY(cudaStreamCreateWithFlags(&mg->capture_stream, cudaStreamNonBlocking));
            cudaGraph_t graph = nullptr;
            const int64_t before = launch_count();
            PANTEA_CUDA_TRY(cudaStreamBeginCapture(mg->capture_stream, cudaStreamCaptureModeThreadLocal));
            int rc = mgpu_step(mg, mg->capture_stream);
            cudaError_t cerr = cudaStreamEndCapture(mg->capture_stream, &graph);
            mg->graph_nodes = (int)(launch_count() - before);
            count_launch(-mg->graph_nodes);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (cerr != cudaSuccess) return fail(PANTEA_ECUDA, std::string("pantea_mgpu_run: graph capture: ") + cudaGetErrorString(cerr));
            cerr = cudaGraphInstantiate(&mg->graph, graph, 0);
            cudaGraphDestroy(graph);
            if (cerr != cudaSuccess) return fail(PANTEA_ECUDA, std::string("pantea_mgpu_run: graph instantiate: ") + cudaGetErrorString(cerr));
            mg->graph_epoch = mg->ws->arg_epoch;
        }
        for (; done < n_steps; ++done) {
            PANTEA_CUDA_TRY(cudaGraphLaunch(mg->graph, st));
            count_launch(mg->graph_nodes);
        }
        return PANTEA_OK;
    }
    for (; done < n_steps; ++done) {
        int rc = mgpu_step(mg, st);
        if (rc) return rc;
    }
    return PANTEA_OK;
}

// copies of the local full-length arrays (only rows with role 2 are authoritative for velocities / forces, rows with
// role >= 1 for positions); any output may be NULL.  *status (HOST): 0, or 1 + rank of a peer that never published a step
// (synchronises `stream`).
int pantea_mgpu_read(pantea_mgpu* mg, void* positions, void* velocities, void* forces, uint8_t* roles, int32_t* status, void* stream) {
    if (!mg) return fail(PANTEA_EINVAL, "pantea_mgpu_read: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = 3 * (size_t)mg->n * esize(mg->dtype);
    if (positions) PANTEA_CUDA_TRY(cudaMemcpyAsync(positions, mg->pos, bytes, cudaMemcpyDeviceToDevice, st));
    if (velocities) PANTEA_CUDA_TRY(cudaMemcpyAsync(velocities, mg->pt_host.vel[mg->rank], bytes, cudaMemcpyDeviceToDevice, st));
    if (forces) PANTEA_CUDA_TRY(cudaMemcpyAsync(forces, mg->pt_host.frc[mg->rank], bytes, cudaMemcpyDeviceToDevice, st));
    if (roles) PANTEA_CUDA_TRY(cudaMemcpyAsync(roles, mg->role, (size_t)mg->n, cudaMemcpyDeviceToDevice, st));
    if (status) {
        int h = 0;
        PANTEA_CUDA_TRY(cudaMemcpyAsync(&h, mg->err, 4, cudaMemcpyDeviceToHost, st));
        PANTEA_CUDA_TRY(cudaStreamSynchronize(st));
        *status = h;
    }
    return PANTEA_OK;
}

int pantea_mgpu_destroy(pantea_mgpu* mg) {
    if (!mg) return PANTEA_OK;
    if (mg->ws && (mg->ws->role == mg->role || mg->world == 1)) { mg->ws->role = nullptr; mg->ws->own_cap = 0; ++mg->ws->arg_epoch; }
    if (mg->graph) cudaGraphExecDestroy(mg->graph);
    if (mg->capture_stream) cudaStreamDestroy(mg->capture_stream);
    for (int r = 0; r < mg->world; ++r)
        if (r != mg->rank && mg->peer_block[r]) cudaIpcCloseMemHandle(mg->peer_block[r]);
    void* ptrs[] = {mg->block, mg->pos, mg->frc_new, mg->role, mg->epoch, mg->done, mg->err, mg->bt_dev, mg->pt_dev};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete mg;
    return PANTEA_OK;
}

}  // extern "C"
