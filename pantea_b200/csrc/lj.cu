// Lennard-Jones pair potential on the library's neighbour rows (second `PotentialInterface` implementation of the
// reference: pantea/simulation/lennard_jones.py:15-123).  Thread per atom over its neighbour row:
//   E = 1/2 sum_i sum_j 4 eps ((sigma/r)^12 - (sigma/r)^6)                     (lennard_jones.py:73-90)
//   "forces"_i = sum_j [-24 eps / r^2 (sigma/r)^6 (2 (sigma/r)^6 - 1)] (r_i - r_j)   (lennard_jones.py:100-123)
// which is +dE/dr_i -- the reference returns the gradient, not the force; kept for parity.
#include "internal.cuh"
#include "math.cuh"

namespace pantea {

template <typename T>
__global__ void lj_kernel(const Rec<T>* __restrict__ rec, const int32_t* __restrict__ nbr, const int32_t* __restrict__ tcount,
                          int cap, int n, const int32_t* __restrict__ struct_of, const double* __restrict__ boxes,
                          double blx, double bly, double blz, int has_box, T sigma, T epsilon, T* __restrict__ e_atom,
                          T* __restrict__ forces) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const Rec<T> ri = rec[slot];
    T lx = (T)blx, ly = (T)bly, lz = (T)blz;
    bool pbc = has_box != 0;
    if (boxes) { const int s = struct_of[slot]; lx = (T)boxes[3 * s]; ly = (T)boxes[3 * s + 1]; lz = (T)boxes[3 * s + 2]; pbc = true; }
    int count = 0;
#pragma unroll
    for (int b = 0; b < kBuckets; ++b) count += tcount[(size_t)slot * kBuckets + b];
    if (count > cap) count = cap;
    const int32_t* row = nbr + (size_t)slot * cap;
    T e = 0, gx = 0, gy = 0, gz = 0;
    for (int q = 0; q < count; ++q) {
        const Rec<T> rj = rec[row[q]];
        T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
        if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
        const T r = norm3_rn(dx, dy, dz);
        const T term = sigma / r;
        const T t2 = term * term, t6 = t2 * t2 * t2;
        e += (T)4 * epsilon * t6 * (t6 - (T)1);
        const T coef = (T)-24 * epsilon / (r * r) * t6 * ((T)2 * t6 - (T)1);
        gx += coef * dx; gy += coef * dy; gz += coef * dz;
    }
    const int out = rec_idx(ri);
    if (e_atom) e_atom[out] = (T)0.5 * e;
    if (forces) { forces[3 * out] = gx; forces[3 * out + 1] = gy; forces[3 * out + 2] = gz; }
}

}  // namespace pantea

using namespace pantea;

extern "C" int pantea_lj_energy_forces(pantea_workspace* ws, double sigma, double epsilon, void* e_atom, void* forces,
                                       void* e_total, void* stream) {
    if (!ws) return fail(PANTEA_EINVAL, "pantea_lj_energy_forces: NULL workspace");
    if (ws->mode == kModeNone) return fail(PANTEA_EINVAL, "pantea_lj_energy_forces: call pantea_neighbor_build first");
    if (ws->skin_active) return fail(PANTEA_EINVAL, "pantea_lj_energy_forces: rows were built with a Verlet skin (set skin = 0)");
    if (!e_atom && !forces && !e_total) return fail(PANTEA_EINVAL, "pantea_lj_energy_forces: all outputs are NULL");
    if (ws->n == 0) return PANTEA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    void* ea = e_atom ? e_atom : (e_total ? ws->md_eatom : nullptr);
    const int blocks = (int)((ws->n + 127) / 128);
    if (ws->dtype == PANTEA_F64)
        lj_kernel<double><<<blocks, 128, 0, st>>>((const Rec<double>*)ws->rec, ws->nbr, ws->nbr_tcount, ws->cap, (int)ws->n,
                                                   ws->struct_of, ws->boxes, ws->box[0], ws->box[1], ws->box[2],
                                                   ws->has_box ? 1 : 0, sigma, epsilon, (double*)ea, (double*)forces);
    else
        lj_kernel<float><<<blocks, 128, 0, st>>>((const Rec<float>*)ws->rec, ws->nbr, ws->nbr_tcount, ws->cap, (int)ws->n,
                                                  ws->struct_of, ws->boxes, ws->box[0], ws->box[1], ws->box[2],
                                                  ws->has_box ? 1 : 0, (float)sigma, (float)epsilon, (float*)ea, (float*)forces);
    PANTEA_LAUNCH_CHECK();
    if (e_total) return reduce_energy(ws, ea, e_total, st);
    return PANTEA_OK;
}
