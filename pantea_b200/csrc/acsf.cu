// Fused per-atom HDNNP kernel for sm_100a:
//   neighbour block staging -> radial (G1/G2) and angular (G3/G9) symmetry functions with analytic
//   central-role gradients -> scaler -> per-element MLP forward/backward -> energy and force.
//
// Replaces the reference's vmap/scan/autodiff pipeline (pantea/descriptors/acsf/acsf.py:163-330,
// descriptors/scaler.py:206-246, models/nn/model.py:53-58, potentials/nnp/energy.py:23-66,
// force.py:16-43).  Work decomposition: WPA warps own one central atom.  The atom's neighbour block
// (d_ij, r_ij, 1/r_ij, fc, fc' per cutoff class; partitioned by neighbour type) is staged once in
// shared memory.  For every neighbour j (warp-uniform, its data broadcast from shared memory) the
// lanes scan the partner neighbours k; pairs surviving the cheap r_jk test are compacted through a
// per-warp shared-memory queue so that the expensive triplet body (sqrt, two exponentials, one
// reciprocal, ~100 FP64 instructions) always runs with full warps.  Partial sums are combined with
// warp shuffles in a fixed order (bitwise reproducible results).  All arithmetic is in T (double or
// float) on the CUDA cores: the path is FP64/FP32-pipe bound, not a GEMM.
#include "internal.cuh"
#include "math.cuh"

namespace pantea {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kAtomsPerBlock = 4;  // WPA == 1 configuration: 4 warps, one atom each
constexpr int kQueue = 64;         // live-pair queue entries per warp

struct BoxArgK {
    double lx, ly, lz;
    int has_box;
};

template <typename T>
struct AtomArgs {
    const Rec<T>* rec;
    const int32_t* nbr;
    const int32_t* tcount;
    int cap;
    const int32_t* slot_of;
    const int32_t* struct_of;
    const double* boxes;
    BoxArgK box;
    int wrap_jk;
    const ElementTable* tables;
    int n_types;
    int element_slot;  // >= 0: apply this element's table to every centre; -1: the atom's own type
    const int32_t* centres;
    int n_work;
    int by_slot;  // 1: work item = cell-sorted slot (energy/force pass); 0: work item = centre list entry
    int own_begin, own_end;
    T* G;
    T* dG;
    int g_stride;
    T* e_atom;
    T* forces;
    unsigned long long* counters;  // optional work counters: [0] pairs, [1] radial-SF evals, [2] triplet-SF evals
    // shared-memory layout (in units of T unless noted)
    int n_cls_max, n_sf_max, n_neurons_max, width_max;
};

template <typename T>
__host__ __device__ inline size_t atom_smem_bytes(int cap, int n_cls, int n_sf, int n_neurons, int width, int wpa) {
    size_t t_elems = (size_t)(5 + 2 * n_cls) * cap   // neighbour block
                     + (size_t)wpa * n_sf * 4        // per-warp partial sums
                     + (size_t)(n_sf + n_neurons)    // layer activations
                     + (size_t)n_neurons             // activation derivatives
                     + 2 * (size_t)(width > n_sf ? width : n_sf);  // back-propagation ping-pong
    size_t bytes = t_elems * sizeof(T);
    bytes = (bytes + 7) & ~size_t(7);
    bytes += (size_t)wpa * kQueue * (sizeof(T) + sizeof(int));  // live-pair queues
    return (bytes + 15) & ~size_t(15);
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

// rarely used, large library routines are kept out of line so that the hot loop stays small
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

template <int WPA>
__device__ __forceinline__ void group_sync() {
    if (WPA == 1) __syncwarp();
    else __syncthreads();
}

// per-atom view of the staged neighbour block
template <typename T>
struct NbrBlock {
    const T *dx, *dy, *dz, *r, *inv, *fc, *dfc;  // fc/dfc already offset to the group's cutoff class
};

// One angular group (same neighbour types, cutoff and kind), MCH members evaluated per triplet.
template <typename T, int WPA, bool GRAD, int MCH>
__device__ __forceinline__ void angular_group(const ElementTable& tab, const AngularGroup& grp, int m0, int mc,
                                              const NbrBlock<T>& nb, int bj, int nj, int bk, int nk, bool wrap_jk, T lx,
                                              T ly, T lz, int lane, int wrank, T* q_r2, int* q_jk, T* my_acc,
                                              unsigned long long& cnt_trip) {
    const bool same = grp.type_j == grp.type_k;
    const int ctype = tab.cls[grp.cls].type;
    const T rc = (T)tab.cls[grp.cls].rc;
    const T inv_rc = (T)1 / rc;
    const T rc2_incl = rc * rc * ((T)1 + (T)8 * (sizeof(T) == 8 ? (T)2.3e-16 : (T)1.2e-7));
    const bool is_g3 = grp.kind == PANTEA_G3;

    T m_neta[MCH], m_lam[MCH], m_zl[MCH], m_pref[MCH], m_zm1[MCH];
    int m_iz[MCH];
#pragma unroll
    for (int m = 0; m < MCH; ++m) {
        const AngularMember mem = tab.members[grp.first + m0 + (m < mc ? m : 0)];
        m_neta[m] = -(T)mem.eta; m_lam[m] = (T)mem.lambda0; m_pref[m] = (T)mem.pref;
        m_zl[m] = (T)(mem.pref * mem.zeta * mem.lambda0); m_zm1[m] = (T)(mem.zeta - 1.0);
        m_iz[m] = mem.izeta;
    }
    T aG[MCH], aX[MCH], aY[MCH], aZ[MCH];
#pragma unroll
    for (int m = 0; m < MCH; ++m) { aG[m] = 0; aX[m] = 0; aY[m] = 0; aZ[m] = 0; }

    // expensive part for one live (j,k) pair
    auto triplet = [&](int jk, T rjk2) {
        const int j = jk & 0xffff, k = jk >> 16;
        const T dxj = nb.dx[j], dyj = nb.dy[j], dzj = nb.dz[j], rj = nb.r[j], ivj = nb.inv[j], fcj = nb.fc[j];
        const T dxk = nb.dx[k], dyk = nb.dy[k], dzk = nb.dz[k], rk = nb.r[k], ivk = nb.inv[k], fck = nb.fc[k];
        T fcjk = (T)1, r2 = rj * rj + rk * rk;
        if (is_g3) { fcjk = cutoff_value_sq<T>(ctype, rjk2, rc, inv_rc); r2 += rjk2; }
        const T ivjk = ivj * ivk;
        const T cost = (dxj * dxk + dyj * dyk + dzj * dzk) * ivjk;
        const T fjk = fcj * fck;
        const T fprod = fjk * fcjk;
        T dfp_j = 0, dfp_k = 0, cj = 0, ck = 0;
        if (GRAD) {
            dfp_j = nb.dfc[j] * fck * fcjk; dfp_k = fcj * nb.dfc[k] * fcjk;
            cj = ivjk - cost * ivj * ivj; ck = ivjk - cost * ivk * ivk;
        }
#pragma unroll
        for (int m = 0; m < MCH; ++m) {
            if (MCH == 1 || m < mc) {
                const T e = fast_exp(m_neta[m] * r2);
                const T bs = (T)1 + m_lam[m] * cost;
                const T pw1 = m_iz[m] == 1 ? (T)1 : (m_iz[m] > 1 ? powi<T>(bs, m_iz[m] - 1) : pow_general<T>(bs, m_zm1[m]));
                const T ang_e = m_pref[m] * pw1 * bs * e;
                aG[m] += ang_e * fprod;
                if (GRAD) {
                    const T Tc = m_zl[m] * pw1 * e * fprod;
                    const T two_neta = (T)2 * m_neta[m];
                    const T Tij = ang_e * (dfp_j + two_neta * rj * fprod);
                    const T Tik = ang_e * (dfp_k + two_neta * rk * fprod);
                    const T Aj = Tc * cj + Tij * ivj, Ak = Tc * ck + Tik * ivk;
                    aX[m] += Aj * dxj + Ak * dxk;
                    aY[m] += Aj * dyj + Ak * dyk;
                    aZ[m] += Aj * dzj + Ak * dzk;
                }
            }
        }
    };

    int qn = 0;
    const int kbase = same ? bj : bk;
    for (int a = wrank; a < nj; a += WPA) {  // neighbour j: uniform within the warp
        const int j = bj + a;
        if (nb.fc[j] == (T)0) continue;      // beyond this group's cutoff
        const T dxj = nb.dx[j], dyj = nb.dy[j], dzj = nb.dz[j];
        for (int kb = same ? a + 1 : 0; kb < nk; kb += 32) {
            const int kk = kb + lane;
            bool live = false;
            T rjk2 = 0;
            const int k = kbase + kk;
            if (kk < nk) {
                T ex = dxj - nb.dx[k], ey = dyj - nb.dy[k], ez = dzj - nb.dz[k];
                if (wrap_jk) { ex = min_image(ex, lx); ey = min_image(ey, ly); ez = min_image(ez, lz); }
                rjk2 = ex * ex + ey * ey + ez * ez;
                live = rjk2 > (T)0 && (!is_g3 || rjk2 < rc2_incl) && nb.fc[k] != (T)0;
            }
            const unsigned mask = __ballot_sync(kFullMask, live);
            if (live) {
                const int pos = qn + __popc(mask & ((1u << lane) - 1u));
                q_jk[pos] = j | (k << 16); q_r2[pos] = rjk2;
            }
            qn += __popc(mask);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                triplet(q_jk[qn + lane], q_r2[qn + lane]);
                cnt_trip += mc;
                __syncwarp();
            }
        }
    }
    if (lane < qn) { triplet(q_jk[lane], q_r2[lane]); cnt_trip += mc; }
    __syncwarp();

#pragma unroll
    for (int m = 0; m < MCH; ++m) {
        if (m < mc) {
            T g = warp_sum(aG[m]), gx = 0, gy = 0, gz = 0;
            if (GRAD) { gx = warp_sum(aX[m]); gy = warp_sum(aY[m]); gz = warp_sum(aZ[m]); }
            if (lane == 0) {
                T* o = my_acc + 4 * tab.members[grp.first + m0 + m].out;
                o[0] = g; o[1] = gx; o[2] = gy; o[3] = gz;
            }
        }
    }
}

template <typename T, int WPA, bool GRAD, int MCH>
__global__ void __launch_bounds__(WPA == 1 ? kAtomsPerBlock * 32 : WPA * 32, (sizeof(T) == 8 && GRAD) ? 4 : 1)
hdnnp_atom_kernel(const AtomArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int atom_in_block = WPA == 1 ? wib : 0;
    const int wrank = WPA == 1 ? 0 : wib;           // rank of this warp among the atom's warps
    const int tid_atom = wrank * 32 + lane;         // thread index within the atom group
    constexpr int S = 32 * WPA;                     // threads per atom
    const int w = blockIdx.x * (WPA == 1 ? kAtomsPerBlock : 1) + atom_in_block;
    if (w >= a.n_work) return;

    // ---- resolve the central atom ---------------------------------------------------------------
    int slot, out_row;
    if (a.by_slot) {
        slot = w;
        const int oi0 = rec_idx(a.rec[slot]);
        if (oi0 < a.own_begin || oi0 >= a.own_end) return;
        out_row = oi0;
    } else {
        const int oi0 = a.centres ? a.centres[w] : w;
        slot = a.slot_of[oi0];
        out_row = w;
    }
    const Rec<T> ri = a.rec[slot];
    const int etype = a.element_slot >= 0 ? a.element_slot : rec_type(ri);
    if (etype >= a.n_types) {  // an atom no element network describes: contributes nothing
        if (tid_atom == 0) {
            if (a.e_atom) a.e_atom[out_row] = (T)0;
            if (a.forces) { a.forces[3 * out_row] = (T)0; a.forces[3 * out_row + 1] = (T)0; a.forces[3 * out_row + 2] = (T)0; }
        }
        return;
    }
    const ElementTable& tab = a.tables[etype];
    const int n_sf = tab.n_sf;

    T lx = (T)a.box.lx, ly = (T)a.box.ly, lz = (T)a.box.lz;
    bool pbc = a.box.has_box != 0;
    if (a.boxes) {
        const int s = a.struct_of[slot];
        lx = (T)a.boxes[3 * s]; ly = (T)a.boxes[3 * s + 1]; lz = (T)a.boxes[3 * s + 2];
        pbc = true;
    }
    const bool wrap_jk = pbc && a.wrap_jk;

    // ---- carve shared memory ---------------------------------------------------------------------
    const int cap = a.cap;
    const size_t per_atom = atom_smem_bytes<T>(cap, a.n_cls_max, a.n_sf_max, a.n_neurons_max, a.width_max, WPA);
    unsigned char* base = smem_raw + (size_t)atom_in_block * per_atom;
    T* sdx = (T*)base;
    T* sdy = sdx + cap;
    T* sdz = sdy + cap;
    T* sr = sdz + cap;
    T* sinv = sr + cap;
    T* sfc = sinv + cap;                       // [n_cls_max][cap]
    T* sdfc = sfc + (size_t)a.n_cls_max * cap; // [n_cls_max][cap]
    T* sacc = sdfc + (size_t)a.n_cls_max * cap;  // [WPA][n_sf_max][4]
    T* sh = sacc + (size_t)WPA * a.n_sf_max * 4; // [n_sf_max + n_neurons_max]
    T* sdact = sh + a.n_sf_max + a.n_neurons_max;
    const int gw = a.width_max > a.n_sf_max ? a.width_max : a.n_sf_max;
    T* sg0 = sdact + a.n_neurons_max;
    T* sg1 = sg0 + gw;
    size_t t_bytes = ((size_t)((sg1 + gw) - sdx) * sizeof(T) + 7) & ~size_t(7);
    T* q_r2 = (T*)(base + t_bytes) + (size_t)wrank * kQueue;
    int* q_jk = (int*)(base + t_bytes + (size_t)WPA * kQueue * sizeof(T)) + (size_t)wrank * kQueue;

    // ---- neighbour segments by type ----------------------------------------------------------------
    int seg[kBuckets + 1];
    {
        const int32_t* tc = a.tcount + (size_t)slot * kBuckets;
        int acc = 0;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) { seg[b] = acc; acc += tc[b]; }
        seg[kBuckets] = acc;
    }
    const int total = seg[kBuckets] < cap ? seg[kBuckets] : cap;
    auto seg_lo = [&](int t) { int v = 0;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) if (b == t) v = seg[b];
        return v < total ? v : total; };
    auto seg_hi = [&](int t) { int v = 0;
#pragma unroll
        for (int b = 0; b < kBuckets; ++b) if (b == t) v = seg[b + 1];
        return v < total ? v : total; };

    // ---- stage the neighbour block -----------------------------------------------------------------
    {
        const int32_t* row = a.nbr + (size_t)slot * cap;
        const int n_cls = tab.n_cls;
        for (int n = tid_atom; n < total; n += S) {
            const Rec<T> rj = a.rec[row[n]];
            T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
            if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
            const T r = norm3_rn(dx, dy, dz);
            sdx[n] = dx; sdy[n] = dy; sdz[n] = dz; sr[n] = r; sinv[n] = (T)1 / r;
            for (int c = 0; c < n_cls; ++c) {
                T fc, dfc;
                cutoff_eval_ool<T>(tab.cls[c].type, r, (T)tab.cls[c].rc, &fc, &dfc);
                sfc[c * cap + n] = fc; sdfc[c * cap + n] = dfc;
            }
        }
    }
    group_sync<WPA>();

    T* my_acc = sacc + (size_t)wrank * a.n_sf_max * 4;
    unsigned long long cnt_rad = 0, cnt_trip = 0;  // per-lane work counters (only summed when requested)

    // ---- radial symmetry functions -----------------------------------------------------------------
    for (int s = 0; s < tab.n_radial; ++s) {
        const RadialSF sf = tab.radial[s];
        const int lo = seg_lo(sf.type_j), hi = seg_hi(sf.type_j);
        const T eta = (T)sf.eta, rs = (T)sf.r_shift;
        const T* fcv = sfc + sf.cls * cap;
        const T* dfcv = sdfc + sf.cls * cap;
        T g = 0, gx = 0, gy = 0, gz = 0;
        for (int n = lo + tid_atom; n < hi; n += S) {
            const T r = sr[n], fc = fcv[n], dfc = dfcv[n];
            T val, dval;
            if (sf.kind == PANTEA_G1) { val = fc; dval = dfc; }
            else {
                const T dr = r - rs, ex = fast_exp(-eta * dr * dr);
                val = ex * fc; dval = ex * (dfc - (T)2 * eta * dr * fc);
            }
            g += val;
            ++cnt_rad;
            if (GRAD) { const T sc = dval * sinv[n]; gx += sc * sdx[n]; gy += sc * sdy[n]; gz += sc * sdz[n]; }
        }
        g = warp_sum(g);
        if (GRAD) { gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz); }
        if (lane == 0) { T* o = my_acc + 4 * sf.out; o[0] = g; o[1] = gx; o[2] = gy; o[3] = gz; }
    }

    // ---- angular symmetry functions ------------------------------------------------------------------
    for (int gi = 0; gi < tab.n_groups; ++gi) {
        const AngularGroup grp = tab.groups[gi];
        const int bj = seg_lo(grp.type_j), nj = seg_hi(grp.type_j) - bj;
        const int bk = seg_lo(grp.type_k), nk = seg_hi(grp.type_k) - bk;
        NbrBlock<T> nb{sdx, sdy, sdz, sr, sinv, sfc + grp.cls * cap, sdfc + grp.cls * cap};
        for (int m0 = 0; m0 < grp.count; m0 += MCH) {
            const int mc = grp.count - m0 < MCH ? grp.count - m0 : MCH;
            angular_group<T, WPA, GRAD, MCH>(tab, grp, m0, mc, nb, bj, nj, bk, nk, wrap_jk, lx, ly, lz, lane, wrank, q_r2,
                                            q_jk, my_acc, cnt_trip);
        }
    }
    if (a.counters) {
        cnt_rad = warp_sum(cnt_rad); cnt_trip = warp_sum(cnt_trip);
        if (lane == 0) {
            if (wrank == 0) atomicAdd(&a.counters[0], (unsigned long long)total);
            atomicAdd(&a.counters[1], cnt_rad);
            atomicAdd(&a.counters[2], cnt_trip);
        }
    }
    group_sync<WPA>();
    if (wrank != 0) return;  // the atom's first warp finishes the job

    // ---- combine the warps' partial sums (fixed order) -------------------------------------------
    if (WPA > 1) {
        for (int e = lane; e < n_sf * 4; e += 32) {
            T v = sacc[e];
#pragma unroll
            for (int q = 1; q < WPA; ++q) v += sacc[(size_t)q * a.n_sf_max * 4 + e];
            sacc[e] = v;
        }
        __syncwarp();
    }
    if (a.G)
        for (int s = lane; s < n_sf; s += 32) a.G[(size_t)out_row * a.g_stride + s] = sacc[4 * s];
    if (GRAD && a.dG)
        for (int e = lane; e < n_sf * 3; e += 32) {
            const int s = e / 3, c = e - 3 * s;
            a.dG[((size_t)out_row * a.g_stride + s) * 3 + c] = sacc[4 * s + 1 + c];
        }
    if (!a.e_atom && !a.forces) return;
    if (tab.n_layers == 0) {
        if (lane == 0) {
            if (a.e_atom) a.e_atom[out_row] = (T)0;
            if (a.forces) { a.forces[3 * out_row] = (T)0; a.forces[3 * out_row + 1] = (T)0; a.forces[3 * out_row + 2] = (T)0; }
        }
        return;
    }

    // ---- scaler + network forward --------------------------------------------------------------------
    for (int s = lane; s < n_sf; s += 32)
        sh[s] = (T)tab.offset[s] + (T)tab.slope[s] * (sacc[4 * s] - (T)tab.shift[s]);
    __syncwarp();
    const int L = tab.n_layers;
    int in_off = 0, out_off = n_sf;
    for (int l = 0; l < L; ++l) {
        const int ni = tab.sizes[l], no = tab.sizes[l + 1];
        const double* W = tab.weights + tab.w_off[l];
        const double* B = W + (size_t)ni * no;
        const int act = tab.acts[l];
        for (int o = lane; o < no; o += 32) {
            T z = (T)0;
            for (int i = 0; i < ni; ++i) z += sh[in_off + i] * (T)W[(size_t)i * no + o];
            z += (T)B[o];
            T y, dy;
            activation_eval_ool<T>(act, z, &y, &dy);
            sh[out_off + o] = y;
            sdact[out_off - n_sf + o] = dy;
        }
        __syncwarp();
        in_off = out_off; out_off += no;
    }
    const T energy = sh[in_off];
    if (lane == 0 && a.e_atom) a.e_atom[out_row] = energy;
    if (!GRAD || !a.forces) return;

    // ---- network backward: dE/dx~ -> dE/dG -> force ---------------------------------------------------
    T* gc = sg0; T* gn = sg1;
    if (lane == 0) gc[0] = (T)1;
    __syncwarp();
    int lay_out = out_off - tab.sizes[L];  // offset (in sh) of the outputs of layer l
    for (int l = L - 1; l >= 0; --l) {
        const int ni = tab.sizes[l], no = tab.sizes[l + 1];
        const double* W = tab.weights + tab.w_off[l];
        const T* da = sdact + (lay_out - n_sf);
        for (int i = lane; i < ni; i += 32) {
            T acc = (T)0;
            for (int o = 0; o < no; ++o) acc += (T)W[(size_t)i * no + o] * (gc[o] * da[o]);
            gn[i] = acc;
        }
        __syncwarp();
        T* tmp = gc; gc = gn; gn = tmp;
        lay_out -= ni;
    }
    T fx = 0, fy = 0, fz = 0;
    for (int s = lane; s < n_sf; s += 32) {
        const T ws_ = gc[s] * (T)tab.slope[s];
        fx -= ws_ * sacc[4 * s + 1]; fy -= ws_ * sacc[4 * s + 2]; fz -= ws_ * sacc[4 * s + 3];
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) { a.forces[3 * out_row] = fx; a.forces[3 * out_row + 1] = fy; a.forces[3 * out_row + 2] = fz; }
}

// ------------------------------------------------------------------------------------------------
// deterministic energy reduction over the owned atoms
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void energy_partial_kernel(const T* __restrict__ e_atom, int begin, int end, int chunk, double* __restrict__ partial) {
    __shared__ double sm[256];
    const int lo = begin + blockIdx.x * chunk, hi = min(lo + chunk, end);
    double v = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += 256) v += (double)e_atom[i];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

template <typename T>
__global__ void energy_final_kernel(const double* __restrict__ partial, int n, T* __restrict__ out) {
    __shared__ double sm[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) v += partial[i];
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (T)sm[0];
}

template <typename T>
static int reduce_energy_typed(pantea_workspace* ws, const T* e_atom, T* e_total, cudaStream_t st) {
    const int begin = (int)ws->own_begin, end = ws->own_end < 0 ? (int)ws->n : (int)ws->own_end;
    const int count = end - begin;
    int chunk = 4096;
    int blocks = (count + chunk - 1) / chunk;
    if (blocks > ws->e_partial_cap) { chunk = (int)((count + ws->e_partial_cap - 1) / ws->e_partial_cap); blocks = (count + chunk - 1) / chunk; }
    if (blocks < 1) blocks = 1;
    energy_partial_kernel<T><<<blocks, 256, 0, st>>>(e_atom, begin, end, chunk, ws->e_partial);
    PANTEA_LAUNCH_CHECK();
    energy_final_kernel<T><<<1, 256, 0, st>>>(ws->e_partial, blocks, e_total);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

int reduce_energy(pantea_workspace* ws, const void* e_atom, void* e_total, cudaStream_t st) {
    if (ws->dtype == PANTEA_F64) return reduce_energy_typed<double>(ws, (const double*)e_atom, (double*)e_total, st);
    return reduce_energy_typed<float>(ws, (const float*)e_atom, (float*)e_total, st);
}

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;

template <typename T, int WPA, bool GRAD, int MCH>
static int launch_cfg(const AtomArgs<T>& args, size_t per_atom, cudaStream_t st) {
    const int apb = WPA == 1 ? kAtomsPerBlock : 1;
    const size_t smem = per_atom * apb;
    auto kern = hdnnp_atom_kernel<T, WPA, GRAD, MCH>;
    static size_t configured = 0;  // per instantiation
    if (smem > configured) {
        if (smem > 227 * 1024) return fail(PANTEA_EINVAL, "neighbour capacity / potential too large for shared memory");
        PANTEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int blocks = (args.n_work + apb - 1) / apb;
    kern<<<blocks, apb * WPA * 32, smem, st>>>(args);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

template <typename T, int WPA, bool GRAD>
static int launch_mch(const AtomArgs<T>& args, size_t per_atom, int max_members, cudaStream_t st) {
    if (max_members <= 1) return launch_cfg<T, WPA, GRAD, 1>(args, per_atom, st);
    if (max_members <= 2) return launch_cfg<T, WPA, GRAD, 2>(args, per_atom, st);
    return launch_cfg<T, WPA, GRAD, 4>(args, per_atom, st);
}

template <typename T>
static int atom_kernel_typed(pantea_workspace* ws, int element_slot, const int32_t* centres, int64_t n_centres, void* G,
                             void* dG, void* e_atom, void* forces, cudaStream_t st) {
    const pantea_potential* pot = ws->pot;
    AtomArgs<T> a{};
    a.rec = (const Rec<T>*)ws->rec; a.nbr = ws->nbr; a.tcount = ws->nbr_tcount; a.cap = ws->cap;
    a.slot_of = ws->slot_of; a.struct_of = ws->struct_of; a.boxes = ws->boxes;
    a.box = BoxArgK{ws->box[0], ws->box[1], ws->box[2], ws->has_box ? 1 : 0};
    double lmin = ws->box[0] < ws->box[1] ? ws->box[0] : ws->box[1];
    if (ws->box[2] < lmin) lmin = ws->box[2];
    a.wrap_jk = ws->boxes ? 1 : (ws->has_box && 0.5 * lmin < 2.0 * ws->rc * (1.0 + 1e-9) ? 1 : 0);
    a.tables = pot->dev; a.n_types = pot->n_elements; a.element_slot = element_slot;
    a.centres = centres;
    const bool energy_pass = (e_atom || forces) && !G && !dG;
    a.by_slot = energy_pass ? 1 : 0;
    a.n_work = energy_pass ? (int)ws->n : (int)n_centres;
    a.own_begin = (int)ws->own_begin; a.own_end = ws->own_end < 0 ? (int)ws->n : (int)ws->own_end;
    a.G = (T*)G; a.dG = (T*)dG; a.e_atom = (T*)e_atom; a.forces = (T*)forces;
    a.counters = ws->counters;
    a.g_stride = element_slot >= 0 ? pot->host[element_slot].n_sf : pot->max_sf;
    a.n_cls_max = pot->max_cls; a.n_sf_max = pot->max_sf > 0 ? pot->max_sf : 1;
    a.n_neurons_max = pot->max_neurons; a.width_max = pot->max_width;
    if (a.n_work == 0) return PANTEA_OK;
    if (g_num_sms == 0) {
        int dev = 0;
        PANTEA_CUDA_TRY(cudaGetDevice(&dev));
        PANTEA_CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const bool grad = dG != nullptr || forces != nullptr;
    const int mm = pot->max_members;
    // few atoms: several warps per atom so that every SM sub-partition has work
    const bool wide = (int64_t)a.n_work < (int64_t)g_num_sms * 64;
    if (wide) {
        const size_t per_atom = atom_smem_bytes<T>(a.cap, a.n_cls_max, a.n_sf_max, a.n_neurons_max, a.width_max, 4);
        return grad ? launch_mch<T, 4, true>(a, per_atom, mm, st) : launch_mch<T, 4, false>(a, per_atom, mm, st);
    }
    const size_t per_atom = atom_smem_bytes<T>(a.cap, a.n_cls_max, a.n_sf_max, a.n_neurons_max, a.width_max, 1);
    return grad ? launch_mch<T, 1, true>(a, per_atom, mm, st) : launch_mch<T, 1, false>(a, per_atom, mm, st);
}

int atom_kernel_launch(pantea_workspace* ws, int element_slot, const int32_t* centres, int64_t n_centres, void* G,
                       void* dG, void* e_atom, void* forces, cudaStream_t st) {
    if (ws->cap > 0xffff) return fail(PANTEA_EINVAL, "max_neighbors must be < 65536");
    if (ws->dtype == PANTEA_F64) return atom_kernel_typed<double>(ws, element_slot, centres, n_centres, G, dG, e_atom, forces, st);
    return atom_kernel_typed<float>(ws, element_slot, centres, n_centres, G, dG, e_atom, forces, st);
}

}  // namespace pantea

using namespace pantea;

extern "C" {

int pantea_acsf_compute(pantea_workspace* ws, int32_t element, const int32_t* centres, int64_t n_centres, void* G,
                        void* dG, void* stream) {
    if (!ws || !ws->pot) return fail(PANTEA_EINVAL, "pantea_acsf_compute: workspace has no potential");
    if (ws->mode == kModeNone) return fail(PANTEA_EINVAL, "pantea_acsf_compute: call pantea_neighbor_build first");
    if (element < 0 || element >= ws->pot->n_elements) return fail(PANTEA_EINVAL, "pantea_acsf_compute: element slot out of range");
    if (!G && !dG) return fail(PANTEA_EINVAL, "pantea_acsf_compute: both outputs are NULL");
    if (!centres) n_centres = ws->n;
    return atom_kernel_launch(ws, element, centres, n_centres, G, dG, nullptr, nullptr, (cudaStream_t)stream);
}

int pantea_energy_forces(pantea_workspace* ws, void* e_atom, void* forces, void* e_total, int32_t force_mode, void* stream) {
    if (!ws || !ws->pot) return fail(PANTEA_EINVAL, "pantea_energy_forces: workspace has no potential");
    if (ws->mode == kModeNone) return fail(PANTEA_EINVAL, "pantea_energy_forces: call pantea_neighbor_build first");
    if (force_mode != PANTEA_FORCE_REFERENCE) return fail(PANTEA_EINVAL, "pantea_energy_forces: unknown force_mode");
    if (!e_atom && !forces && !e_total) return fail(PANTEA_EINVAL, "pantea_energy_forces: all outputs are NULL");
    if (ws->n == 0) return PANTEA_OK;
    void* ea = e_atom ? e_atom : (e_total ? ws->md_eatom : nullptr);
    int rc = atom_kernel_launch(ws, -1, nullptr, 0, nullptr, nullptr, ea, forces, (cudaStream_t)stream);
    if (rc != PANTEA_OK) return rc;
    if (e_total) return reduce_energy(ws, ea, e_total, (cudaStream_t)stream);
    return PANTEA_OK;
}

}  // extern "C"
