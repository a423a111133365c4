// Halo exchange helpers of the brick-decomposed MD (the reference is single-process and has
// no counterpart -- the exchanged quantity is the `positions` argument of its jitted kernels, acsf.py:210-228).
//
// Between two rebuilds of the ghost lists the messages have a fixed shape, so one step needs exactly one launch here:
//   * gather the owned atoms every neighbouring brick needs (fixed send list) into the contiguous send buffer that the
//     all-to-all reads, and
//   * check the Verlet criterion of the ghost shell: an owned atom that moved more than `limit` (= skin / 2, minimum
//     image) since the lists were built raises a sticky device flag, read by the host only at the next rebuild.
// The reverse direction (full-force mode) adds the ghost contributions that came back onto the owned atoms in the
// order of the send list, one thread per owned atom walking its (sorted) occurrences: no atomics, reproducible.
#include "internal.cuh"
#include "math.cuh"

namespace pantea {

template <typename T>
__global__ void halo_pack_kernel(const T* __restrict__ pos, const int64_t* __restrict__ send_idx, int64_t n_send,
                                 T* __restrict__ send_buf, const T* __restrict__ pos_ref, int64_t n_own, double lx,
                                 double ly, double lz, int has_box, double limit2, int32_t* __restrict__ violated) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_send) {
        const int64_t s = send_idx[i];
        send_buf[3 * i] = pos[3 * s]; send_buf[3 * i + 1] = pos[3 * s + 1]; send_buf[3 * i + 2] = pos[3 * s + 2];
    }
    if (pos_ref && i < n_own) {
        double dx = (double)pos[3 * i] - (double)pos_ref[3 * i], dy = (double)pos[3 * i + 1] - (double)pos_ref[3 * i + 1],
               dz = (double)pos[3 * i + 2] - (double)pos_ref[3 * i + 2];
        if (has_box) {  // wrapped coordinates: the displacement is the nearest image
            dx -= lx * rint(dx / lx); dy -= ly * rint(dy / ly); dz -= lz * rint(dz / lz);
        }
        if (dx * dx + dy * dy + dz * dz > limit2) *violated = 1;
    }
}

// out[idx] += buf over the occurrences of each owned atom; `order` sorts the send list by atom (stable), `first[a]`
// .. `first[a+1]` are atom a's occurrences in `order`
template <typename T>
__global__ void halo_unpack_add_kernel(T* __restrict__ out, const T* __restrict__ buf, const int64_t* __restrict__ order,
                                       const int64_t* __restrict__ first, int64_t n_own) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_own) return;
    T x = out[3 * a], y = out[3 * a + 1], z = out[3 * a + 2];
    for (int64_t o = first[a]; o < first[a + 1]; ++o) {
        const int64_t e = order[o];
        x += buf[3 * e]; y += buf[3 * e + 1]; z += buf[3 * e + 2];
    }
    out[3 * a] = x; out[3 * a + 1] = y; out[3 * a + 2] = z;
}

}  // namespace pantea

using namespace pantea;

extern "C" {

int pantea_halo_pack(const void* positions, const int64_t* send_idx, int64_t n_send, void* send_buf, const void* pos_ref,
                     int64_t n_own, const double* box, double limit, int32_t* violated, int32_t dtype, void* stream) {
    if (dtype != PANTEA_F64 && dtype != PANTEA_F32) return fail(PANTEA_EINVAL, "pantea_halo_pack: dtype must be PANTEA_F64 or PANTEA_F32");
    if (n_send < 0 || n_own < 0) return fail(PANTEA_EINVAL, "pantea_halo_pack: negative count");
    if (n_send > 0 && (!positions || !send_idx || !send_buf)) return fail(PANTEA_EINVAL, "pantea_halo_pack: NULL array");
    if (pos_ref && (!violated || !positions)) return fail(PANTEA_EINVAL, "pantea_halo_pack: pos_ref needs positions and a flag");
    const int64_t work = pos_ref ? (n_send > n_own ? n_send : n_own) : n_send;
    if (work == 0) return PANTEA_OK;
    const double lx = box ? box[0] : 0.0, ly = box ? box[1] : 0.0, lz = box ? box[2] : 0.0;
    const int blocks = (int)((work + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PANTEA_F64)
        halo_pack_kernel<double><<<blocks, 256, 0, st>>>((const double*)positions, send_idx, n_send, (double*)send_buf,
                                                         (const double*)pos_ref, n_own, lx, ly, lz, box ? 1 : 0,
                                                         limit * limit, violated);
    else
        halo_pack_kernel<float><<<blocks, 256, 0, st>>>((const float*)positions, send_idx, n_send, (float*)send_buf,
                                                        (const float*)pos_ref, n_own, lx, ly, lz, box ? 1 : 0,
                                                        limit * limit, violated);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int pantea_halo_unpack_add(void* out, const void* recv_buf, const int64_t* order, const int64_t* first, int64_t n_own,
                           int32_t dtype, void* stream) {
    if (dtype != PANTEA_F64 && dtype != PANTEA_F32) return fail(PANTEA_EINVAL, "pantea_halo_unpack_add: dtype must be PANTEA_F64 or PANTEA_F32");
    if (n_own < 0) return fail(PANTEA_EINVAL, "pantea_halo_unpack_add: negative count");
    if (n_own == 0) return PANTEA_OK;
    if (!out || !recv_buf || !order || !first) return fail(PANTEA_EINVAL, "pantea_halo_unpack_add: NULL array");
    const int blocks = (int)((n_own + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == PANTEA_F64)
        halo_unpack_add_kernel<double><<<blocks, 256, 0, st>>>((double*)out, (const double*)recv_buf, order, first, n_own);
    else
        halo_unpack_add_kernel<float><<<blocks, 256, 0, st>>>((float*)out, (const float*)recv_buf, order, first, n_own);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

}  // extern "C"
