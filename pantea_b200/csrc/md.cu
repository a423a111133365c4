// Velocity-Verlet / Berendsen kernels and the device-resident MD loop.
//
// Replaces reference pantea/simulation/molecular_dynamics.py:16-77 (no mass in the integrator),
// atoms/box.py:123-126 (floored-remainder wrap), simulation/system.py:20-29 (kinetic energy,
// temperature) and simulation/thermostat.py:12-22.  These are streaming, HBM-bound kernels
// (~150 B per atom and step); the MD loop replays them together with the neighbour build and the
// fused energy/force kernel as one CUDA graph per step, with no host synchronisation.
#include <cstring>

#include "internal.cuh"
#include "math.cuh"

namespace pantea {

template <typename T> __device__ __forceinline__ T t_fmod(T a, T b);
template <> __device__ __forceinline__ double t_fmod<double>(double a, double b) { return fmod(a, b); }
template <> __device__ __forceinline__ float t_fmod<float>(float a, float b) { return fmodf(a, b); }

// floored remainder, like jnp.remainder (reference box.py:123-126)
template <typename T>
__device__ __forceinline__ T wrap_coord(T x, T box) {
    T m = t_fmod<T>(x, box);
    if (m != (T)0 && ((m < (T)0) != (box < (T)0))) m = add_rn(m, box);
    return m;
}

struct Box3 {
    double l[3];
    int has_box;
};

// `mass` (nullable) selects the mass-scaled extension: the acceleration F/m takes the place of F (the
// reference integrator itself carries no mass)
template <typename T>
__global__ void md_positions_kernel(T* __restrict__ pos, const T* __restrict__ vel, const T* __restrict__ frc,
                                    const T* __restrict__ mass, int64_t begin3, int64_t end3, Box3 box, T dt) {
    int64_t e = begin3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= end3) return;
    const T a = mass ? frc[e] / mass[e / 3] : frc[e];
    // x + v*dt + 0.5*F*dt*dt evaluated left to right without contraction (molecular_dynamics.py:20)
    T x = add_rn(add_rn(pos[e], mul_rn(vel[e], dt)), mul_rn(mul_rn(mul_rn((T)0.5, a), dt), dt));
    if (box.has_box) x = wrap_coord<T>(x, (T)box.l[e % 3]);
    pos[e] = x;
}

template <typename T>
__global__ void md_velocities_kernel(T* __restrict__ vel, T* __restrict__ frc, const T* __restrict__ frc_new,
                                     const T* __restrict__ mass, int64_t begin3, int64_t end3, T dt) {
    int64_t e = begin3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= end3) return;
    const T fn = frc_new[e];
    T f0 = frc[e], f1 = fn;
    if (mass) { const T m = mass[e / 3]; f0 = f0 / m; f1 = f1 / m; }
    vel[e] = add_rn(vel[e], mul_rn(mul_rn((T)0.5, add_rn(f0, f1)), dt));  // molecular_dynamics.py:30
    frc[e] = fn;
}

constexpr int kKeChunk = 2048;  // atoms per partial sum

template <typename T>
__global__ void md_ke_partial_kernel(const T* __restrict__ vel, const T* __restrict__ mass, int64_t begin, int64_t end,
                                     double* __restrict__ partial) {
    __shared__ double sm[256];
    const int64_t lo = begin + (int64_t)blockIdx.x * kKeChunk;
    const int64_t hi = lo + kKeChunk < end ? lo + kKeChunk : end;
    double v = 0.0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
        const double m = (double)mass[i];
        const double vx = (double)vel[3 * i], vy = (double)vel[3 * i + 1], vz = (double)vel[3 * i + 2];
        v += m * vx * vx + m * vy * vy + m * vz * vz;
    }
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void md_ke_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sm[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v += partial[i];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = 0.5 * sm[0];  // system.py:21-22
}

__device__ __forceinline__ double berendsen_factor(double ke, int64_t n_total, double dt, double tau, double t0, double kb) {
    const double temp = 2.0 * ke / (3.0 * (double)n_total * kb);          // system.py:25-29
    return 1.0 / sqrt(1.0 + (dt / tau) * (temp / t0 - 1.0));             // thermostat.py:16-21
}

template <typename T>
__global__ void md_rescale_kernel(T* __restrict__ vel, int64_t begin3, int64_t end3, const double* __restrict__ ke,
                                  int64_t n_total, double dt, double tau, double t0, double kb) {
    int64_t e = begin3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= end3) return;
    const T s = (T)berendsen_factor(ke[0], n_total, dt, tau, t0, kb);
    vel[e] = mul_rn(vel[e], s);
}

// scalars[2*step] = E_pot, scalars[2*step+1] = E_kin (after the thermostat), step counter advanced on device
template <typename T>
__global__ void md_record_kernel(const T* __restrict__ e_total, const double* __restrict__ ke, int* __restrict__ counter,
                                 double* __restrict__ scalars, int64_t n_total, double dt, double tau, double t0, double kb) {
    const int step = counter[0];
    double k = ke[0];
    if (tau > 0.0) { const double s = berendsen_factor(k, n_total, dt, tau, t0, kb); k *= s * s; }
    scalars[2 * step] = (double)e_total[0];
    scalars[2 * step + 1] = k;
    counter[0] = step + 1;
}

static Box3 make_box(const double* box) {
    Box3 b{{0, 0, 0}, 0};
    if (box) { b.l[0] = box[0]; b.l[1] = box[1]; b.l[2] = box[2]; b.has_box = 1; }
    return b;
}

template <typename T>
static int kinetic_energy_typed(const void* vel, const void* mass, int64_t begin, int64_t end, double* partial,
                                int64_t partial_cap, double* out, cudaStream_t st) {
    const int64_t count = end - begin;
    int blocks = (int)((count + kKeChunk - 1) / kKeChunk);
    if (blocks < 1) blocks = 1;
    if (blocks > partial_cap) return fail(PANTEA_EINVAL, "kinetic energy: too many atoms for the reduction scratch");
    md_ke_partial_kernel<T><<<blocks, 256, 0, st>>>((const T*)vel, (const T*)mass, begin, end, partial);
    PANTEA_LAUNCH_CHECK();
    md_ke_final_kernel<<<1, 256, 0, st>>>(partial, blocks, out);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

}  // namespace pantea

using namespace pantea;

static double* g_ke_scratch_dev[64] = {nullptr};  // reduction scratch of the stand-alone entry point, per device
static const int64_t kKeScratchCap = 1 << 17;

extern "C" {

int pantea_md_update_positions_mass(void* positions, const void* velocities, const void* forces, const void* masses,
                                    int64_t begin, int64_t end, const double* box, double dt, int32_t dtype, void* stream) {
    if (!positions || !velocities || !forces) return fail(PANTEA_EINVAL, "pantea_md_update_positions: NULL argument");
    if (dtype != PANTEA_F64 && dtype != PANTEA_F32) return fail(PANTEA_EINVAL, "pantea_md_update_positions: dtype must be PANTEA_F64 or PANTEA_F32");
    if (end <= begin) return PANTEA_OK;
    const int64_t n3 = 3 * (end - begin);
    const int blocks = (int)((n3 + 255) / 256);
    Box3 b = make_box(box);
    if (dtype == PANTEA_F64)
        md_positions_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)positions, (const double*)velocities,
                                                                             (const double*)forces, (const double*)masses,
                                                                             3 * begin, 3 * end, b, dt);
    else
        md_positions_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)positions, (const float*)velocities,
                                                                            (const float*)forces, (const float*)masses,
                                                                            3 * begin, 3 * end, b, (float)dt);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int pantea_md_update_positions(void* positions, const void* velocities, const void* forces, int64_t begin, int64_t end,
                               const double* box, double dt, int32_t dtype, void* stream) {
    return pantea_md_update_positions_mass(positions, velocities, forces, nullptr, begin, end, box, dt, dtype, stream);
}

int pantea_md_update_velocities_mass(void* velocities, void* forces, const void* new_forces, const void* masses,
                                     int64_t begin, int64_t end, double dt, int32_t dtype, void* stream) {
    if (!velocities || !forces || !new_forces) return fail(PANTEA_EINVAL, "pantea_md_update_velocities: NULL argument");
    if (dtype != PANTEA_F64 && dtype != PANTEA_F32) return fail(PANTEA_EINVAL, "pantea_md_update_velocities: dtype must be PANTEA_F64 or PANTEA_F32");
    if (end <= begin) return PANTEA_OK;
    const int64_t n3 = 3 * (end - begin);
    const int blocks = (int)((n3 + 255) / 256);
    if (dtype == PANTEA_F64)
        md_velocities_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)velocities, (double*)forces,
                                                                              (const double*)new_forces, (const double*)masses,
                                                                              3 * begin, 3 * end, dt);
    else
        md_velocities_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)velocities, (float*)forces,
                                                                             (const float*)new_forces, (const float*)masses,
                                                                             3 * begin, 3 * end, (float)dt);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int pantea_md_update_velocities(void* velocities, void* forces, const void* new_forces, int64_t begin, int64_t end,
                                double dt, int32_t dtype, void* stream) {
    return pantea_md_update_velocities_mass(velocities, forces, new_forces, nullptr, begin, end, dt, dtype, stream);
}

int pantea_md_kinetic_energy(const void* velocities, const void* masses, int64_t begin, int64_t end, double* ke_out,
                             int32_t dtype, void* stream) {
    if (!velocities || !masses || !ke_out) return fail(PANTEA_EINVAL, "pantea_md_kinetic_energy: NULL argument");
    int dev = 0;
    PANTEA_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(PANTEA_EINVAL, "device index out of range");
    if (!g_ke_scratch_dev[dev]) PANTEA_CUDA_TRY(cudaMalloc((void**)&g_ke_scratch_dev[dev], 8 * kKeScratchCap));
    double* g_ke_scratch = g_ke_scratch_dev[dev];
    if (dtype == PANTEA_F64)
        return kinetic_energy_typed<double>(velocities, masses, begin, end, g_ke_scratch, kKeScratchCap, ke_out, (cudaStream_t)stream);
    return kinetic_energy_typed<float>(velocities, masses, begin, end, g_ke_scratch, kKeScratchCap, ke_out, (cudaStream_t)stream);
}

int pantea_md_rescale_velocities(void* velocities, int64_t begin, int64_t end, const double* ke, int64_t n_total, double dt,
                                 double tau, double t_target, double kb, int32_t dtype, void* stream) {
    if (!velocities || !ke) return fail(PANTEA_EINVAL, "pantea_md_rescale_velocities: NULL argument");
    if (!(tau > 0.0) || !(t_target > 0.0)) return fail(PANTEA_EINVAL, "pantea_md_rescale_velocities: tau and t_target must be positive");
    if (end <= begin) return PANTEA_OK;
    const int64_t n3 = 3 * (end - begin);
    const int blocks = (int)((n3 + 255) / 256);
    if (dtype == PANTEA_F64)
        md_rescale_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)velocities, 3 * begin, 3 * end, ke, n_total, dt, tau, t_target, kb);
    else
        md_rescale_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)velocities, 3 * begin, 3 * end, ke, n_total, dt, tau, t_target, kb);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

// one velocity-Verlet step on `st` (reference molecular_dynamics.py:57-77)
static int md_step(pantea_workspace* ws, void* pos, void* vel, void* frc, const void* mass, const int32_t* types, int64_t n,
                   const double* box, const pantea_md_params* p, double* scalars, cudaStream_t st) {
    const void* m_int = p->mass_scaled ? mass : nullptr;  // mass-scaled extension (the reference integrator has no mass)
    int rc = pantea_md_update_positions_mass(pos, vel, frc, m_int, 0, n, box, p->dt, ws->dtype, st);
    if (rc) return rc;
    rc = neighbor_build_impl(ws, pos, types, n, box, nullptr, nullptr, 1, ws->pot->rc_max, st);
    if (rc) return rc;
    const bool record = p->record && scalars;
    rc = atom_kernel_launch(ws, -1, nullptr, 0, nullptr, nullptr, record ? ws->md_eatom : nullptr, ws->md_forces, st,
                            p->force_mode);
    if (rc) return rc;
    rc = pantea_md_update_velocities_mass(vel, frc, ws->md_forces, m_int, 0, n, p->dt, ws->dtype, st);
    if (rc) return rc;
    const bool thermo = p->tau > 0.0;
    if (thermo || record) {
        if (ws->dtype == PANTEA_F64) rc = kinetic_energy_typed<double>(vel, mass, 0, n, ws->e_partial, ws->e_partial_cap, ws->md_ke, st);
        else rc = kinetic_energy_typed<float>(vel, mass, 0, n, ws->e_partial, ws->e_partial_cap, ws->md_ke, st);
        if (rc) return rc;
    }
    if (thermo) {
        rc = pantea_md_rescale_velocities(vel, 0, n, ws->md_ke, n, p->dt, p->tau, p->t_target, p->kb, ws->dtype, st);
        if (rc) return rc;
    }
    if (record) {
        // E_pot total into md_ke[1] (as T), then the record kernel
        void* etot = (void*)(ws->md_ke + 1);
        rc = reduce_energy(ws, ws->md_eatom, etot, st);
        if (rc) return rc;
        int* counter = (int*)(ws->md_ke + 2);
        if (ws->dtype == PANTEA_F64)
            md_record_kernel<double><<<1, 1, 0, st>>>((const double*)etot, ws->md_ke, counter, scalars, n, p->dt, p->tau, p->t_target, p->kb);
        else
            md_record_kernel<float><<<1, 1, 0, st>>>((const float*)etot, ws->md_ke, counter, scalars, n, p->dt, p->tau, p->t_target, p->kb);
        PANTEA_LAUNCH_CHECK();
    }
    return PANTEA_OK;
}

int pantea_md_run(pantea_workspace* ws, void* positions, void* velocities, void* forces, const void* masses,
                  const int32_t* types, int64_t n_atoms, const double* box, int64_t n_steps, const pantea_md_params* params,
                  double* scalars, void* stream) {
    if (!ws || !ws->pot) return fail(PANTEA_EINVAL, "pantea_md_run: workspace has no potential");
    if (!positions || !velocities || !forces || !masses || !types || !params) return fail(PANTEA_EINVAL, "pantea_md_run: NULL argument");
    if (n_atoms < 1 || n_atoms > ws->max_atoms) return fail(PANTEA_EINVAL, "pantea_md_run: n_atoms out of range");
    if ((n_atoms + kKeChunk - 1) / kKeChunk > ws->e_partial_cap) return fail(PANTEA_EINVAL, "pantea_md_run: too many atoms for the reduction scratch");
    if (params->force_mode != PANTEA_FORCE_REFERENCE && params->force_mode != PANTEA_FORCE_FULL)
        return fail(PANTEA_EINVAL, "pantea_md_run: unknown force_mode");
    if (n_steps <= 0) return PANTEA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool record = params->record && scalars;
    if (record) PANTEA_CUDA_TRY(cudaMemsetAsync(ws->md_ke + 2, 0, 8, st));
    int64_t done = 0;
    // first step eagerly: performs every lazy allocation / attribute setting outside of graph capture
    int rc = md_step(ws, positions, velocities, forces, masses, types, n_atoms, box, params, scalars, st);
    if (rc) return rc;
    done = 1;
    if (params->use_graph && n_steps - done >= 2) {
        pantea_workspace::GraphKey key;
        key.pos = positions; key.vel = velocities; key.frc = forces; key.mass = masses; key.types = types;
        key.scalars = record ? scalars : nullptr; key.n = n_atoms; key.dt = params->dt; key.tau = params->tau;
        key.t0 = params->t_target; key.kb = params->kb; key.record = record ? 1 : 0; key.has_box = box ? 1 : 0;
        key.mass_scaled = params->mass_scaled ? 1 : 0; key.force_mode = params->force_mode;
        for (int k = 0; k < 3; ++k) key.box[k] = box ? box[k] : 0.0;
        key.epoch = ws->arg_epoch;  // (after the eager step: its lazy allocations are in)
        if (!(ws->md_graph && key == ws->md_key)) {
            if (ws->md_graph) { cudaGraphExecDestroy(ws->md_graph); ws->md_graph = nullptr; }
            cudaGraph_t graph = nullptr;
            const int64_t before = launch_count();
            // capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture
            if (!ws->capture_stream) PANTEA_CUDA_TRY(cudaStreamCreateWithFlags(&ws->capture_stream, cudaStreamNonBlocking));
            cudaStream_t cs = ws->capture_stream;
            PANTEA_CUDA_TRY(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
            rc = md_step(ws, positions, velocities, forces, masses, types, n_atoms, box, params, scalars, cs);
            cudaError_t cerr = cudaStreamEndCapture(cs, &graph);
            ws->md_graph_nodes = (int)(launch_count() - before);
            count_launch(-ws->md_graph_nodes);  // captured, not executed
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (cerr != cudaSuccess) return fail(PANTEA_ECUDA, std::string("pantea_md_run: graph capture: ") + cudaGetErrorString(cerr));
            cerr = cudaGraphInstantiate(&ws->md_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (cerr != cudaSuccess) return fail(PANTEA_ECUDA, std::string("pantea_md_run: graph instantiate: ") + cudaGetErrorString(cerr));
            ws->md_key = key;
        }
        for (; done < n_steps; ++done) {
            PANTEA_CUDA_TRY(cudaGraphLaunch(ws->md_graph, st));
            count_launch(ws->md_graph_nodes);
        }
        return PANTEA_OK;
    }
    for (; done < n_steps; ++done) {
        rc = md_step(ws, positions, velocities, forces, masses, types, n_atoms, box, params, scalars, st);
        if (rc) return rc;
    }
    return PANTEA_OK;
}

}  // extern "C"
