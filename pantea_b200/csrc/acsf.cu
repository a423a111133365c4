// HDNNP symmetry-function / network / force kernels for sm_100a.
//
// Replaces the reference's vmap/scan/autodiff pipeline (pantea/descriptors/acsf/acsf.py:163-330,
// descriptors/scaler.py:206-246, models/nn/model.py:53-58, potentials/nnp/energy.py:23-66,
// force.py:16-43) with two launches per evaluation:
//
//  1. `pair_filter_kernel` (FP32 / integer pipes, high occupancy).  One warp per central atom stages a float4
//     copy of its neighbour vectors in shared memory and, for every angular group (type_j, type_k, cutoff, kind),
//     scans the (j,k) neighbour pairs -- j warp-uniform, lanes over k, 64 partners per iteration -- keeping those
//     whose r_jk^2 lies below a slightly inclusive single-precision bound.  Survivors are ballot-compacted into a
//     per-atom pair list in global memory (coalesced 4-byte stores, deterministic order).
//  2. `hdnnp_eval_kernel` (FP64 / FP32 arithmetic pipe).  WPA warps per central atom stage the full-precision
//     neighbour block (records d_ij, r_ij, 1/r_ij, fc, fc' per cutoff class) in shared memory,
//     evaluate the radial functions (lanes over neighbours) and walk the pair lists flat: every lane evaluates NU
//     independent triplets per iteration (r_jk recomputed exactly; the reference's r_jk > 0 and r_jk < rc tests are
//     applied here), so the ~100-FP64-instruction triplet body runs with full warps and NU independent dependency
//     chains.  Warp-shuffle reductions in a fixed order (bitwise reproducible), then scaler -> MLP forward /
//     backward (lanes over neurons) -> E_i and F_i = -sum_s dE_i/dG_s dG_s/dr_i (central-role gradient: no scatter).
//
// All arithmetic is on the CUDA cores: the path is FP64/FP32-pipe bound, not a GEMM.
#include <atomic>
#include <cstdlib>

#include "acsf_common.cuh"

namespace pantea {

// ------------------------------------------------------------------------------------------------
// 1. pair pre-filter
// ------------------------------------------------------------------------------------------------
// One sweep of the pair filter: neighbours j in [j_lo, j_hi) of the bucket starting at `js` against the lane's partner.
// ORDER: also require j < k (the chunk's own neighbours in a same-type group); CLS: test j's cutoff-class bit (only
// groups whose cutoff is shorter than the list radius); WRAP: minimum image on r_jk.  Two neighbours per iteration;
// the staging array has one spare entry so that the odd tail can be read unconditionally.
// true when the neighbours at row positions n1 and n2 of the centre in `slot` sit at exactly the same position as seen
// from the centre (d_ij == d_ik in every component: the reference's r_jk is then 0 and it drops the triplet,
// acsf.py:316-325).  Only called when the binning flagged coincident atoms and the staged single-precision vectors agree.
template <typename T>
__device__ __noinline__ bool coincident_neighbours(const Rec<T>* __restrict__ rec, const int32_t* __restrict__ row, int slot,
                                                   int n1, int n2, T lx, T ly, T lz, bool pbc) {
    const Rec<T> ri = rec[slot];
    const Rec<T> r1 = rec[row[n1]], r2 = rec[row[n2]];
    T d1[3] = {sub_rn(ri.x, r1.x), sub_rn(ri.y, r1.y), sub_rn(ri.z, r1.z)};
    T d2[3] = {sub_rn(ri.x, r2.x), sub_rn(ri.y, r2.y), sub_rn(ri.z, r2.z)};
    if (pbc) {
        d1[0] = min_image(d1[0], lx); d1[1] = min_image(d1[1], ly); d1[2] = min_image(d1[2], lz);
        d2[0] = min_image(d2[0], lx); d2[1] = min_image(d2[1], ly); d2[2] = min_image(d2[2], lz);
    }
    return d1[0] == d2[0] && d1[1] == d2[1] && d1[2] == d2[2];
}

template <typename T>
struct FilterSweep {
    int32_t* list;
    int pair_cap, cls_bit, k_hi, kk;
    unsigned lt_mask;
    float rc2f, flx, fly, flz;
    float4 fk;
    bool k_ok;
    // exact coincidence test, only when the binning saw atoms at the same position
    bool exact, pbc;
    const Rec<T>* rec;
    const int32_t* row;
    int slot, k_pos;
    T lx, ly, lz;
    template <bool ORDER, bool CLS, bool WRAP>
    __device__ __forceinline__ int run(const float4* __restrict__ sf4, int js, int j_lo, int j_hi, int off) const {
        for (int aj = j_lo; aj < j_hi; aj += 2) {
            const float4 fj[2] = {sf4[js + aj], sf4[js + aj + 1]};
            bool live[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float ex = fj[c].x - fk.x, ey = fj[c].y - fk.y, ez = fj[c].z - fk.z;
                if (WRAP) { ex = min_image(ex, flx); ey = min_image(ey, fly); ez = min_image(ez, flz); }
                const float d2 = ex * ex + ey * ey + ez * ez;
                live[c] = k_ok && d2 < rc2f;
                if (exact && live[c] && fj[c].x == fk.x && fj[c].y == fk.y && fj[c].z == fk.z && js + aj + c != k_pos)
                    live[c] = !coincident_neighbours<T>(rec, row, slot, js + aj + c, k_pos, lx, ly, lz, pbc);
                if (CLS) live[c] = live[c] && (__float_as_int(fj[c].w) & cls_bit);
                if (ORDER) live[c] = live[c] && aj + c < kk;
            }
            live[1] = live[1] && aj + 1 < j_hi;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const unsigned mask = __ballot_sync(kFullMask, live[c]);
                const int pos = off + __popc(mask & lt_mask);
                if (live[c] && pos < pair_cap) list[pos] = (js + aj + c) | k_hi;
                off += __popc(mask);
            }
        }
        return off;
    }
};

template <typename T>
__global__ void __launch_bounds__(kFilterWarps * 32) pair_filter_kernel(const AtomArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.filter_guard && *a.filter_guard == 0) return;  // rows unchanged since the lists were written
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kFilterWarps + wib;
    if (w >= a.n_work) return;
    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    int32_t* off_out = a.pair_off + (size_t)w * (a.max_groups + 1);
    if (etype >= a.n_types) {
        if (lane == 0) off_out[0] = 0;
        return;
    }
    const ElementTable& tab = a.tables[etype];
    float4* sf4 = (float4*)smem_raw + (size_t)wib * (a.scap + 1);  // + 1: see FilterSweep

    Segments sg;
    sg.load(a.tcount + (size_t)slot * kBuckets, a.scap);
    if (sg.seg[kBuckets] > a.scap && lane == 0) atomicMax(&a.flags[1], sg.seg[kBuckets]);

    T lx, ly, lz;
    bool pbc;
    item_box(a, slot, lx, ly, lz, pbc);
    const bool wrap_jk = pbc && a.wrap_jk;
    const float flx = (float)lx, fly = (float)ly, flz = (float)lz;

    // stage (dx, dy, dz, cutoff-class bit mask): differences formed in T, then rounded to float
    {
        const Rec<T> ri = a.rec[slot];
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        const int n_cls = tab.n_cls;
        for (int n = lane; n < sg.total; n += 32) {
            const Rec<T> rj = a.rec[row[n]];
            T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
            if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
            const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
            const float r2 = fx * fx + fy * fy + fz * fz;
            int bits = 0;
            for (int c = 0; c < n_cls; ++c) {
                const float rcf = (float)tab.cls[c].rc + a.skin;
                bits |= (r2 < rcf * rcf * 1.0001f + 1e-4f) ? (1 << c) : 0;  // inclusive: exact test in the evaluation
            }
            sf4[n] = make_float4(fx, fy, fz, __int_as_float(bits));
        }
    }
    __syncwarp();

    int32_t* list = a.pairs + (size_t)w * a.pair_cap;
    int pair_cap = a.pair_cap;
    asm volatile("" : "+l"(list), "+r"(pair_cap));  // keep both in registers: the store below is one IMAD.WIDE + STG
    const unsigned lt_mask = (1u << lane) - 1u;
    int off = 0;  // running number of pairs (may exceed pair_cap: then only counted)
    for (int gi = 0; gi < tab.n_groups; ++gi) {
        if (lane == 0) off_out[gi] = off < a.pair_cap ? off : a.pair_cap;
        const AngularGroup grp = tab.groups[gi];
        const int bj = sg.lo(grp.type_j), nj = sg.hi(grp.type_j) - bj;
        const int bk = sg.lo(grp.type_k), nk = sg.hi(grp.type_k) - bk;
        const bool same = grp.type_j == grp.type_k;
        const float rcf = (float)tab.cls[grp.cls].rc + a.skin;
        const float rc2f = grp.kind == PANTEA_G3 ? rcf * rcf * 1.0001f + 1e-4f : 3.0e38f;
        // every lane keeps one partner k in registers while the warp sweeps over the neighbours j (broadcast reads).
        // Same-type groups take the unordered pairs j < k: neighbours before the chunk pair with all of its lanes, the
        // chunk's own neighbours need the order test.  Mixed groups sweep the bucket that leaves fewer idle lanes.
        // The pair order (k chunk, j, k) is fixed, hence the evaluation's summation order is deterministic.
        FilterSweep<T> sw;
        sw.exact = a.dup_always || (a.dup_flag && *a.dup_flag != 0); sw.pbc = pbc; sw.rec = a.rec; sw.row = a.nbr + (size_t)slot * a.cap;
        sw.slot = slot; sw.lx = lx; sw.ly = ly; sw.lz = lz;
        sw.list = list; sw.pair_cap = pair_cap; sw.lt_mask = lt_mask; sw.rc2f = rc2f; sw.cls_bit = 1 << grp.cls;
        sw.flx = flx; sw.fly = fly; sw.flz = flz;
        const bool swap = !same && nj * ((nk + 31) >> 5) > nk * ((nj + 31) >> 5);
        const int js = swap ? bk : bj, njs = swap ? nk : nj;  // swept bucket
        const int ks = swap ? bj : bk, nks = swap ? nj : nk;  // lane-resident bucket
        const int variant = (tab.cls[grp.cls].rc + (double)a.skin < a.rc_list ? 1 : 0) | (wrap_jk ? 2 : 0);
        for (int k0 = 0; k0 < nks; k0 += 32) {
            const int kk = k0 + lane;
            sw.fk = sf4[ks + (kk < nks ? kk : 0)];
            sw.k_ok = kk < nks && (__float_as_int(sw.fk.w) & sw.cls_bit);
            sw.k_hi = (ks + kk) << 16;
            sw.k_pos = ks + (kk < nks ? kk : 0);
            sw.kk = kk;
            const int j_full = same ? min(njs, k0) : njs, j_diag = same ? min(njs, k0 + 31) : njs;
            switch (variant) {
                case 0: off = sw.run<false, false, false>(sf4, js, 0, j_full, off); off = sw.run<true, false, false>(sf4, js, j_full, j_diag, off); break;
                case 1: off = sw.run<false, true, false>(sf4, js, 0, j_full, off); off = sw.run<true, true, false>(sf4, js, j_full, j_diag, off); break;
                case 2: off = sw.run<false, false, true>(sf4, js, 0, j_full, off); off = sw.run<true, false, true>(sf4, js, j_full, j_diag, off); break;
                default: off = sw.run<false, true, true>(sf4, js, 0, j_full, off); off = sw.run<true, true, true>(sf4, js, j_full, j_diag, off); break;
            }
        }
    }
    if (lane == 0) {
        off_out[tab.n_groups] = off < a.pair_cap ? off : a.pair_cap;
        atomicMax(&a.flags[2], off);
    }
}

// ------------------------------------------------------------------------------------------------
// 2. evaluation
// ------------------------------------------------------------------------------------------------
// staged neighbour block: array of records [dx, dy, dz, r, 1/r, (fc, fc') per cutoff class]; the record length
// 5 + 2 n_cls is odd, so consecutive as well as scattered records spread over the shared-memory banks, and one
// address computation per neighbour serves all of its fields
template <typename T>
struct NbrBlock {
    const T* rec;
    int stride, fco;  // record length; offset of (fc, fc') of the group's cutoff class
};

// One angular group (same neighbour types, cutoff and kind), members [m0, m0 + mc), flat over the pair list.
// FAST: compile-time specialisation for the common RuNNer setting -- G3, tanhu cutoff, integer zeta >= 1 for every member
// -- whose triplet body is straight-line code; otherwise kind / cutoff / zeta are runtime (warp-uniform) branches.
// LEAN (implies FAST, no counters): additionally zeta == 1 for every member and no minimum image on r_jk; list slots
// past the end are staged as `pad_entry` = (total, total), the all-zero record behind the atom's last neighbour (u = 0,
// r = 0, fc = 0, fc'/fc = 0: every term of the triplet body is then exactly zero whatever r_jk evaluates to), so the
// loop carries no validity mask, no power loop and no wrap branch.
template <typename T, int WPA, bool GRAD, int MCH, bool FAST, bool COUNT, bool LEAN = false>
__device__ __forceinline__ void angular_group(const ElementTable& tab, const AngularGroup& grp, int m0, int mc,
                                              const NbrBlock<T>& nb, const int32_t* __restrict__ list, int count,
                                              bool wrap_jk, T lx, T ly, T lz, int lane, int tid_atom, T* my_acc,
                                              const T* __restrict__ etab, int* stage, unsigned long long& cnt_trip,
                                              int pad_entry = 0) {
    constexpr int NU = kNU;
    constexpr int S = 32 * WPA;
    const int ctype = tab.cls[grp.cls].type;
    const T rc = (T)tab.cls[grp.cls].rc;
    const T inv_rc = (T)1 / rc, m2_inv_rc = (T)-2 / rc;
    const bool is_g3 = FAST ? true : grp.kind == PANTEA_G3;

    T m_neta[MCH], m_2neta[MCH], m_lam[MCH], m_zl[MCH], m_pref[MCH], m_zm1[MCH];
    int m_iz[MCH];
#pragma unroll
    for (int m = 0; m < MCH; ++m) {
        const AngularMember mem = tab.members[grp.first + m0 + (m < mc ? m : 0)];
        m_neta[m] = -(T)mem.eta; m_2neta[m] = (T)(-2.0 * mem.eta); m_lam[m] = (T)mem.lambda0; m_pref[m] = (T)mem.pref;
        m_zl[m] = (T)(mem.pref * mem.zeta * mem.lambda0); m_zm1[m] = (T)(mem.zeta - 1.0);
        m_iz[m] = mem.izeta;
    }
    T aG[MCH], aX[MCH], aY[MCH], aZ[MCH];
#pragma unroll
    for (int m = 0; m < MCH; ++m) { aG[m] = 0; aX[m] = 0; aY[m] = 0; aZ[m] = 0; }

    // software pipeline over the pair list (it lives in HBM/L2): every lane copies its own entries of the next
    // kStageIters iterations into a shared-memory ring with cp.async while the current chunk is evaluated. The copies
    // are tracked by cp.async groups, not by the register scoreboard, so a consumer never waits on a younger load.
    constexpr int CH = kStageIters;
    const int n_iter = (count + S * NU - 1) / (S * NU);
    const int n_chunks = (n_iter + CH - 1) / CH;
    auto stage_chunk = [&](int chunk) {  // entries past the end re-read entry 0: every lane always holds a real pair
        int* dst = stage + (chunk & 1) * (CH * NU * 32) + lane;
        int e = chunk * (CH * NU * S) + tid_atom;
#pragma unroll
        for (int i = 0; i < CH * NU; ++i, e += S) {
            if (LEAN) {  // the lane reads back only what it wrote itself: a plain store needs no fence here
                if (e < count) cp_async4(dst + i * 32, list + e);
                else dst[i * 32] = pad_entry;
            } else {
                cp_async4(dst + i * 32, list + (e < count ? e : 0));
            }
        }
        cp_async_commit();
    };
    if (n_chunks > 0) stage_chunk(0);
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        if (chunk + 1 < n_chunks) {
            stage_chunk(chunk + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        const int* src = stage + (chunk & 1) * (CH * NU * 32) + lane;
        int e0 = chunk * (CH * NU * S) + tid_atom;
        const int n_in = min(CH, n_iter - chunk * CH);
        for (int i = 0; i < n_in; ++i, src += NU * 32, e0 += NU * S) {
            bool valid[NU];
            // neighbour records: [u_x, u_y, u_z, r, 1/r, (fc, fc'/fc) per cutoff class], u = d / r
            T uxj[NU], uyj[NU], uzj[NU], rj[NU], ivj[NU], fcj[NU], uxk[NU], uyk[NU], uzk[NU], rk[NU], ivk[NU], fck[NU];
            T qj[NU], qk[NU], r2[NU], rjk2[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int jk = src[u * 32];
                valid[u] = LEAN ? true : e0 + u * S < count;
                const T* pj = nb.rec + (jk & 0xffff) * nb.stride;
                const T* pk = nb.rec + (jk >> 16) * nb.stride;
                uxj[u] = pj[0]; uyj[u] = pj[1]; uzj[u] = pj[2]; rj[u] = pj[3]; ivj[u] = pj[4]; fcj[u] = pj[nb.fco];
                uxk[u] = pk[0]; uyk[u] = pk[1]; uzk[u] = pk[2]; rk[u] = pk[3]; ivk[u] = pk[4]; fck[u] = pk[nb.fco];
                qj[u] = GRAD ? pj[nb.fco + 1] : (T)0; qk[u] = GRAD ? pk[nb.fco + 1] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {  // r_jk = |pbc(d_ij - d_ik)| (reference acsf.py:316-320)
                T ex = uxj[u] * rj[u] - uxk[u] * rk[u], ey = uyj[u] * rj[u] - uyk[u] * rk[u], ez = uzj[u] * rj[u] - uzk[u] * rk[u];
                if (!LEAN && wrap_jk) { ex = min_image(ex, lx); ey = min_image(ey, ly); ez = min_image(ez, lz); }
                rjk2[u] = ex * ex + ey * ey + ez * ez;
                r2[u] = rj[u] * rj[u] + rk[u] * rk[u];
            }
            T fcjk[NU];
            if (FAST) {
                // coincident neighbours (r_jk == 0, excluded by acsf.py:325) give rsqrt(0) -> NaN in rjk and t; the
                // comparison below is then false and the select discards them
                T rjk[NU], t[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) rjk[u] = fast_sqrt_loop(rjk2[u]);
#pragma unroll
                for (int u = 0; u < NU; ++u) t[u] = fast_tanh_half_tab<T>(fma(rjk[u], m2_inv_rc, (T)2), etab);
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    fcjk[u] = (valid[u] && rjk[u] < rc) ? t[u] * t[u] * t[u] : (T)0;
                    r2[u] += rjk2[u];
                }
            } else {
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    valid[u] = valid[u] && rjk2[u] > (T)0;  // k == j / coincident atoms excluded (acsf.py:325)
                    fcjk[u] = (T)1;
                    if (is_g3) { fcjk[u] = cutoff_value_sq<T>(ctype, valid[u] ? rjk2[u] : (T)1, rc, inv_rc); r2[u] += rjk2[u]; }
                    if (!valid[u]) fcjk[u] = (T)0;
                }
            }
            // with P = fc_j fc_k fc_jk, A = pref (1 + lambda cos)^zeta exp(-eta r2):   G += A P  and
            //   dG/dr_i -= sum over j of [A P (q_j - 2 eta r_j) + T_c (1/r_k - cos / r_j)] u_j,   q = fc'/fc,
            //   T_c = pref zeta lambda (1 + lambda cos)^(zeta-1) exp(-eta r2) P     (sign folded into the caller)
            T cost[NU], fprod[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                cost[u] = uxj[u] * uxk[u] + uyj[u] * uyk[u] + uzj[u] * uzk[u];
                fprod[u] = fcj[u] * fck[u] * fcjk[u];
            }
#pragma unroll
            for (int m = 0; m < MCH; ++m) {
                if (MCH == 1 || m < mc) {
                    T e[NU];
#pragma unroll
                    for (int u = 0; u < NU; ++u) e[u] = fast_exp_tab<true>(m_neta[m] * r2[u], etab);
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        const T bs = (T)1 + m_lam[m] * cost[u];
                        T pw1 = (T)1;
                        if (!LEAN && m_iz[m] != 1)  // warp-uniform; FAST groups only hold integer zeta >= 1
                            pw1 = (FAST || m_iz[m] > 1) ? powi<T>(bs, m_iz[m] - 1) : pow_general<T>(bs, m_zm1[m]);
                        const T ep = e[u] * fprod[u] * pw1;
                        // zeta == 0: x^0 = 1 also at x = 0 (the reference's pow), where bs * bs^-1 would be NaN
                        const T ap = (!FAST && m_iz[m] == 0) ? m_pref[m] * e[u] * fprod[u] : m_pref[m] * bs * ep;
                        aG[m] += ap;
                        if (GRAD) {
                            const T Tc = (!FAST && m_iz[m] == 0) ? (T)0 : m_zl[m] * ep;
                            const T Bj = Tc * (ivk[u] - cost[u] * ivj[u]) + ap * (qj[u] + m_2neta[m] * rj[u]);
                            const T Bk = Tc * (ivj[u] - cost[u] * ivk[u]) + ap * (qk[u] + m_2neta[m] * rk[u]);
                            aX[m] = fma(Bk, uxk[u], fma(Bj, uxj[u], aX[m]));
                            aY[m] = fma(Bk, uyk[u], fma(Bj, uyj[u], aY[m]));
                            aZ[m] = fma(Bk, uzk[u], fma(Bj, uzj[u], aZ[m]));
                        }
                    }
                }
            }
            if (COUNT && cnt_trip != ~0ull) {
#pragma unroll
                for (int u = 0; u < NU; ++u) cnt_trip += (valid[u] && fcjk[u] != (T)0) ? (unsigned long long)mc : 0ull;
            }
        }
    }

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
__global__ void __launch_bounds__(kEvalWarps * 32, (sizeof(T) == 8 && GRAD) ? PANTEA_EVAL_MINBLOCKS : 1)
hdnnp_eval_kernel(const AtomArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int APB = kEvalWarps / WPA;           // atoms per block
    const int atom_in_block = wib / WPA;
    const int wrank = wib % WPA;                    // rank of this warp among the atom's warps
    const int tid_atom = wrank * 32 + lane;         // thread index within the atom group
    constexpr int S = 32 * WPA;                     // threads per atom
    __shared__ T s_etab[64];                        // 2^(i/64) for the table-driven exponential
    exp2_table_fill(s_etab, threadIdx.x, blockDim.x);
    __shared__ int s_stage[4][2 * kStageIters * kNU * 32];  // per-warp pair-list staging ring (128 threads per block)
    __syncthreads();
    const int w = blockIdx.x * APB + atom_in_block;
    if (w >= a.n_work) return;

    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    if (etype >= a.n_types) return;  // an atom no element network describes (the network kernel writes zeros)
    const ElementTable& tab = a.tables[etype];
    const int n_sf = tab.n_sf;
    const Rec<T> ri = a.rec[slot];

    T lx, ly, lz;
    bool pbc;
    item_box(a, slot, lx, ly, lz, pbc);
    const bool wrap_jk = pbc && a.wrap_jk;

    // ---- carve shared memory ---------------------------------------------------------------------
    const int cap = a.scap;
    const size_t per_atom = eval_smem_bytes<T>(cap, a.n_cls_max, a.n_sf_max, a.n_neurons_max, a.width_max, WPA);
    const int stride = 5 + 2 * a.n_cls_max;
    T* snb = (T*)(smem_raw + (size_t)atom_in_block * per_atom);  // [cap + 1][stride] neighbour records
    T* sacc = snb + (size_t)stride * (cap + 1);                  // [WPA][n_sf_max][4]

    Segments sg;
    sg.load(a.tcount + (size_t)slot * kBuckets, cap);
    const int total = sg.total;
    // a row longer than the staged capacity is cut here: report it (radial-only potentials launch no pair filter)
    if (sg.seg[kBuckets] > cap && lane == 0 && wrank == 0) atomicMax(&a.flags[1], sg.seg[kBuckets]);

    // ---- stage the neighbour block (two neighbours per thread in flight: gather latency) ------------
    {
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        const int n_cls = tab.n_cls;
        for (int n0 = tid_atom; n0 < total; n0 += 2 * S) {
            const int n1 = n0 + S;
            const bool has1 = n1 < total;
            const int i0 = row[n0], i1 = row[has1 ? n1 : n0];
            const Rec<T> ra = a.rec[i0], rb = a.rec[i1];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !has1) break;
                const Rec<T>& rj = u == 0 ? ra : rb;
                const int n = u == 0 ? n0 : n1;
                T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
                if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
                // r and 1/r from one reciprocal-root seed (the exact-rounding rule only matters for the neighbour
                // predicate, which neighbor_rows_kernel has already applied)
                const T r2n = dx * dx + dy * dy + dz * dz;
                const T iv = fast_rsqrt(r2n), r = r2n * iv;
                T* p = snb + (size_t)n * stride;
                p[0] = dx * iv; p[1] = dy * iv; p[2] = dz * iv; p[3] = r; p[4] = iv;
                for (int c = 0; c < n_cls; ++c) {
                    T fc, q;  // cutoff value and its logarithmic derivative fc'/fc
                    const int ct = tab.cls[c].type;
                    const T rcc = (T)tab.cls[c].rc;
                    if (ct == PANTEA_CUT_TANHU) {  // inline fast path for the common cutoff
                        const T irc = fast_rcp(rcc);
                        const T t = fast_tanh_pos<T>((T)1 - r * irc);
                        const bool in = r < rcc && t > (T)0;
                        fc = in ? t * t * t : (T)0;
                        q = in ? (T)-3 * irc * ((T)1 - t * t) * fast_rcp(t) : (T)0;
                    } else {
                        T dfc;
                        cutoff_eval_ool<T>(ct, r, rcc, &fc, &dfc);
                        q = fc != (T)0 ? dfc / fc : (T)0;
                    }
                    p[5 + 2 * c] = fc; p[6 + 2 * c] = q;
                }
            }
        }
    }
    if (tid_atom < stride) snb[(size_t)total * stride + tid_atom] = (T)0;  // padding record of the LEAN triplet loop
    group_sync<WPA>(atom_in_block);

    T* my_acc = sacc + (size_t)wrank * a.n_sf_max * 4;
    unsigned long long cnt_rad = 0, cnt_trip = a.counters ? 0ull : ~0ull;  // work counters (~0: disabled)

    // ---- radial symmetry functions -----------------------------------------------------------------
    for (int s = 0; s < tab.n_radial; ++s) {
        const RadialSF sf = tab.radial[s];
        const int lo = sg.lo(sf.type_j), hi = sg.hi(sf.type_j);
        const T eta = (T)sf.eta, rs = (T)sf.r_shift;
        const int fco = 5 + 2 * sf.cls;
        T g = 0, gx = 0, gy = 0, gz = 0;
        for (int n = lo + tid_atom; n < hi; n += S) {
            const T* p = snb + (size_t)n * stride;
            const T r = p[3], fc = p[fco], q = p[fco + 1];
            T val, dval;
            if (sf.kind == PANTEA_G1) { val = fc; dval = fc * q; }
            else {
                const T dr = r - rs, ex = fast_exp(-eta * dr * dr);
                val = ex * fc; dval = val * (q - (T)2 * eta * dr);
            }
            g += val;
            ++cnt_rad;
            if (GRAD) { gx += dval * p[0]; gy += dval * p[1]; gz += dval * p[2]; }
        }
        g = warp_sum(g);
        if (GRAD) { gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz); }
        if (lane == 0) { T* o = my_acc + 4 * sf.out; o[0] = g; o[1] = gx; o[2] = gy; o[3] = gz; }
    }

    // ---- angular symmetry functions: flat walk over the pre-filtered pair lists ---------------------
    {
        const int32_t* offs = a.pair_off + (size_t)w * (a.max_groups + 1);
        const int32_t* lists = a.pairs + (size_t)w * a.pair_cap;
        for (int gi = 0; gi < tab.n_groups; ++gi) {
            const AngularGroup grp = tab.groups[gi];
            const int lo = offs[gi], count = offs[gi + 1] - lo;
            NbrBlock<T> nb{snb, stride, 5 + 2 * grp.cls};
            bool fast = tab.cls[grp.cls].type == PANTEA_CUT_TANHU && grp.kind == PANTEA_G3;
            bool zeta1 = true;
            for (int m = 0; m < grp.count; ++m) {
                fast = fast && tab.members[grp.first + m].izeta >= 1;
                zeta1 = zeta1 && tab.members[grp.first + m].izeta == 1;
            }
            for (int m0 = 0; m0 < grp.count; m0 += MCH) {
                const int mc = grp.count - m0 < MCH ? grp.count - m0 : MCH;
                // the work counters (bench / roofline pass only) are compiled out of the common fast variant
                if (fast && zeta1 && !wrap_jk && cnt_trip == ~0ull)
                    angular_group<T, WPA, GRAD, MCH, true, false, true>(tab, grp, m0, mc, nb, lists + lo, count, wrap_jk, lx, ly, lz,
                                                                        lane, tid_atom, my_acc, s_etab, s_stage[threadIdx.x >> 5], cnt_trip,
                                                                        total | (total << 16));
                else if (fast && cnt_trip == ~0ull)
                    angular_group<T, WPA, GRAD, MCH, true, false>(tab, grp, m0, mc, nb, lists + lo, count, wrap_jk, lx, ly, lz,
                                                                  lane, tid_atom, my_acc, s_etab, s_stage[threadIdx.x >> 5], cnt_trip);
                else if (fast)
                    angular_group<T, WPA, GRAD, MCH, true, true>(tab, grp, m0, mc, nb, lists + lo, count, wrap_jk, lx, ly, lz,
                                                                 lane, tid_atom, my_acc, s_etab, s_stage[threadIdx.x >> 5], cnt_trip);
                else
                    angular_group<T, WPA, GRAD, MCH, false, true>(tab, grp, m0, mc, nb, lists + lo, count, wrap_jk, lx, ly, lz,
                                                                  lane, tid_atom, my_acc, s_etab, s_stage[threadIdx.x >> 5], cnt_trip);
            }
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
    group_sync<WPA>(atom_in_block);
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
    // hand the summed descriptor (value + central gradient) to the network kernel
    if (a.gbuf)
        for (int e = lane; e < n_sf * 4; e += 32) a.gbuf[(size_t)w * a.n_sf_max * 4 + e] = sacc[e];
}

// ------------------------------------------------------------------------------------------------
// 3. scaler + per-element network forward / backward + force: one THREAD per atom (the networks are tiny:
//    3-5-5-1 ... 30-25-25-1; a warp per atom would leave most lanes idle).  Per-thread scratch lives in shared
//    memory, strided by thread.  Replaces scaler.py:206-246, model.py:53-58, energy.py:36-39, force.py:16-43.
// ------------------------------------------------------------------------------------------------
constexpr int kMlpThreads = 64;

template <typename T>
__global__ void __launch_bounds__(kMlpThreads) mlp_force_kernel(const AtomArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = blockIdx.x * kMlpThreads + threadIdx.x;
    if (w >= a.n_work) return;
    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    T energy = (T)0, fx = (T)0, fy = (T)0, fz = (T)0;
    if (etype < a.n_types && a.tables[etype].n_layers > 0) {
        const ElementTable& tab = a.tables[etype];
        const int n_sf = tab.n_sf, L = tab.n_layers;
        const T* g = a.gbuf + (size_t)w * a.n_sf_max * 4;
        // per-thread scratch, element i of thread t at [i * kMlpThreads + t]
        T* sh = (T*)smem_raw + threadIdx.x;                                       // [n_sf_max + n_neurons_max]
        T* sdact = sh + (size_t)(a.n_sf_max + a.n_neurons_max) * kMlpThreads;       // [n_neurons_max]
        const int gw = a.width_max > a.n_sf_max ? a.width_max : a.n_sf_max;
        T* gc = sdact + (size_t)a.n_neurons_max * kMlpThreads;                      // [gw]
        T* gn = gc + (size_t)gw * kMlpThreads;                                      // [gw]
        for (int s = 0; s < n_sf; ++s)
            sh[s * kMlpThreads] = (T)tab.offset[s] + (T)tab.slope[s] * (g[4 * s] - (T)tab.shift[s]);
        int in_off = 0, out_off = n_sf;
        for (int l = 0; l < L; ++l) {
            const int ni = tab.sizes[l], no = tab.sizes[l + 1];
            const double* W = tab.weights + tab.w_off[l];
            const double* B = W + (size_t)ni * no;
            const int act = tab.acts[l];
            // narrow layers (RuNNer: 5 ... 8 neurons): all outputs accumulate together -- `no` independent chains of length
            // `ni` with the layer's weight rows read as consecutive words, instead of one chain of ni x no dependent
            // loads and adds; same summation order per neuron (i ascending, then the bias)
            constexpr int NW = 8;
            T zz[NW];
            const bool narrow = no <= NW;
            if (narrow) {
#pragma unroll
                for (int o = 0; o < NW; ++o) zz[o] = (T)0;
                for (int i = 0; i < ni; ++i) {
                    const T hi = sh[(in_off + i) * kMlpThreads];
#pragma unroll
                    for (int o = 0; o < NW; ++o) if (o < no) zz[o] += hi * (T)W[(size_t)i * no + o];
                }
#pragma unroll
                for (int o = 0; o < NW; ++o) if (o < no) zz[o] += (T)B[o];
            }
            for (int o = 0; o < no; ++o) {
                T z = (T)0;
                if (narrow) {
#pragma unroll
                    for (int q = 0; q < NW; ++q) if (q == o) z = zz[q];
                } else {
                    for (int i = 0; i < ni; ++i) z += sh[(in_off + i) * kMlpThreads] * (T)W[(size_t)i * no + o];
                    z += (T)B[o];
                }
                T y, dy;
                if (act == PANTEA_ACT_TANH) {
                    // inline tanh(z) = sign(z) (1 - 2 / (exp(2|z|) + 1)): the library routine is ~150 instructions of a
                    // kernel that is a serial chain per atom (latency-bound when a rank owns ~10^4 atoms); abs. error 1e-16
                    const T az = z < (T)0 ? -z : z;
                    const T t = az < (T)20 ? (T)1 - (T)2 * fast_rcp(fast_exp((T)2 * az) + (T)1) : (T)1;
                    y = z < (T)0 ? -t : t; dy = (T)1 - t * t;
                } else {
                    activation_eval_ool<T>(act, z, &y, &dy);
                }
                sh[(out_off + o) * kMlpThreads] = y;
                sdact[(out_off - n_sf + o) * kMlpThreads] = dy;
            }
            in_off = out_off; out_off += no;
        }
        energy = sh[in_off * kMlpThreads];
        if (a.forces || a.wbuf) {
            gc[0] = (T)1;
            int lay_out = out_off - tab.sizes[L];  // offset (in sh) of the outputs of layer l
            for (int l = L - 1; l >= 0; --l) {
                const int ni = tab.sizes[l], no = tab.sizes[l + 1];
                const double* W = tab.weights + tab.w_off[l];
                const T* da = sdact + (size_t)(lay_out - n_sf) * kMlpThreads;
                if (no <= 8) {  // the products gc * da once, then one short chain per input, four inputs in flight
                    T t[8];
#pragma unroll
                    for (int o = 0; o < 8; ++o) t[o] = o < no ? gc[o * kMlpThreads] * da[o * kMlpThreads] : (T)0;
#pragma unroll 4
                    for (int i = 0; i < ni; ++i) {
                        T acc = (T)0;
#pragma unroll
                        for (int o = 0; o < 8; ++o) if (o < no) acc += (T)W[(size_t)i * no + o] * t[o];
                        gn[i * kMlpThreads] = acc;
                    }
                } else {
                    for (int i = 0; i < ni; ++i) {
                        T acc = (T)0;
                        for (int o = 0; o < no; ++o) acc += (T)W[(size_t)i * no + o] * (gc[o * kMlpThreads] * da[o * kMlpThreads]);
                        gn[i * kMlpThreads] = acc;
                    }
                }
                T* tmp = gc; gc = gn; gn = tmp;
                lay_out -= ni;
            }
            for (int s = 0; s < n_sf; ++s) {
                const T ws_ = gc[s * kMlpThreads] * (T)tab.slope[s];
                if (a.wbuf) a.wbuf[(size_t)w * a.n_sf_max + s] = ws_;
                fx -= ws_ * g[4 * s + 1]; fy -= ws_ * g[4 * s + 2]; fz -= ws_ * g[4 * s + 3];
            }
        }
    }
    if (a.e_atom) a.e_atom[out_row] = energy;
    if (a.forces) { a.forces[3 * out_row] = fx; a.forces[3 * out_row + 1] = fy; a.forces[3 * out_row + 2] = fz; }
}

// ------------------------------------------------------------------------------------------------
// 4. full-force mode (PANTEA_FORCE_FULL; SURVEY.md 8(f)-4, not a reference mode): F = -dE/dr of the total energy.
//    With w_s = dE_i/dG_is from the network kernel, every centre i contributes g_in = sum_s w_s dG_is/d(d_in),
//    d_in = r_i - r_n, to itself (F_i -= g_in) and to its neighbour n (F_n += g_in).  Unlike the central-role
//    derivative (acsf.py:215-228) the r_jk leg of G3 does not cancel:
//      dT/dd_ij = T_c (u_k - c u_j)/r_j + T_j u_j + T_jk u_jk,   dT/dd_ik = T_c (u_j - c u_k)/r_k + T_k u_k - T_jk u_jk
//    (SURVEY.md Appendix A).  One warp per centre walks the radial functions and the same pre-filtered pair lists
//    as the evaluation, accumulates g_in per neighbour in shared memory (shared atomics) and issues one global
//    atomic per neighbour and component.  Generic over kinds / cutoffs / zeta (runtime, warp-uniform branches).
// ------------------------------------------------------------------------------------------------
constexpr int kFullWarps = 4;

template <typename T>
__host__ __device__ inline size_t full_smem_bytes(int cap, int n_cls, int n_sf) {
    const size_t t_elems = (size_t)(5 + 2 * n_cls) * cap + (size_t)3 * cap + (size_t)n_sf;
    return ((t_elems * sizeof(T) + (size_t)cap * sizeof(int)) + 15) & ~size_t(15);
}

template <typename T>
__global__ void __launch_bounds__(kFullWarps * 32) full_force_kernel(const AtomArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * kFullWarps + wib;
    if (w >= a.n_work) return;
    int slot, out_row, etype;
    if (!resolve_item(a, w, slot, out_row, etype)) return;
    if (etype >= a.n_types || a.tables[etype].n_layers == 0) return;
    const ElementTable& tab = a.tables[etype];
    const Rec<T> ri = a.rec[slot];

    T lx, ly, lz;
    bool pbc;
    item_box(a, slot, lx, ly, lz, pbc);
    const bool wrap_jk = pbc && a.wrap_jk;

    const int cap = a.scap;
    const int stride = 5 + 2 * a.n_cls_max;
    unsigned char* base = smem_raw + (size_t)wib * full_smem_bytes<T>(cap, a.n_cls_max, a.n_sf_max);
    T* snb = (T*)base;                          // [cap][stride]: u_x, u_y, u_z, r, 1/r, (fc, fc') per cutoff class
    T* sacc = snb + (size_t)stride * cap;       // [cap][3] force on each neighbour from this centre
    T* sw = sacc + (size_t)3 * cap;             // [n_sf_max] dE_i/dG_is
    int* sidx = (int*)(sw + a.n_sf_max);        // [cap] original atom index of each neighbour

    Segments sg;
    sg.load(a.tcount + (size_t)slot * kBuckets, cap);
    const int total = sg.total;
    if (sg.seg[kBuckets] > cap && lane == 0) atomicMax(&a.flags[1], sg.seg[kBuckets]);
    {
        const int32_t* row = a.nbr + (size_t)slot * a.cap;
        const int n_cls = tab.n_cls;
        for (int n = lane; n < total; n += 32) {
            const Rec<T> rj = a.rec[row[n]];
            T dx = sub_rn(ri.x, rj.x), dy = sub_rn(ri.y, rj.y), dz = sub_rn(ri.z, rj.z);
            if (pbc) { dx = min_image(dx, lx); dy = min_image(dy, ly); dz = min_image(dz, lz); }
            const T r = t_sqrt<T>(dx * dx + dy * dy + dz * dz), iv = (T)1 / r;
            T* p = snb + (size_t)n * stride;
            p[0] = dx * iv; p[1] = dy * iv; p[2] = dz * iv; p[3] = r; p[4] = iv;
            for (int c = 0; c < n_cls; ++c) {
                T fc, dfc;
                cutoff_eval_ool<T>(tab.cls[c].type, r, (T)tab.cls[c].rc, &fc, &dfc);
                p[5 + 2 * c] = fc; p[6 + 2 * c] = dfc;
            }
            sacc[3 * n] = (T)0; sacc[3 * n + 1] = (T)0; sacc[3 * n + 2] = (T)0;
            sidx[n] = rec_idx(rj);
        }
        for (int s = lane; s < tab.n_sf; s += 32) sw[s] = a.wbuf[(size_t)w * a.n_sf_max + s];
    }
    __syncwarp();

    // ---- radial: g_in = sum_s w_s g_s'(r_in) u_in ; every lane owns its neighbours, no atomics ------------
    for (int s = 0; s < tab.n_radial; ++s) {
        const RadialSF sf = tab.radial[s];
        const int lo = sg.lo(sf.type_j), hi = sg.hi(sf.type_j);
        const T eta = (T)sf.eta, rs = (T)sf.r_shift, ws_ = sw[sf.out];
        const int fco = 5 + 2 * sf.cls;
        for (int n = lo + lane; n < hi; n += 32) {
            const T* p = snb + (size_t)n * stride;
            const T r = p[3], fc = p[fco], dfc = p[fco + 1];
            T dval;
            if (sf.kind == PANTEA_G1) dval = dfc;
            else {
                const T dr = r - rs, ex = t_exp<T>(-eta * dr * dr);
                dval = ex * (dfc - (T)2 * eta * dr * fc);
            }
            const T c = ws_ * dval;
            sacc[3 * n] += c * p[0]; sacc[3 * n + 1] += c * p[1]; sacc[3 * n + 2] += c * p[2];
        }
        __syncwarp();
    }

    // ---- angular: walk the pair lists, one triplet per lane and iteration ---------------------------------
    {
        const int32_t* offs = a.pair_off + (size_t)w * (a.max_groups + 1);
        const int32_t* lists = a.pairs + (size_t)w * a.pair_cap;
        for (int gi = 0; gi < tab.n_groups; ++gi) {
            const AngularGroup grp = tab.groups[gi];
            const int lo = offs[gi], count = offs[gi + 1] - lo;
            const int fco = 5 + 2 * grp.cls;
            const int ctype = tab.cls[grp.cls].type;
            const T rc = (T)tab.cls[grp.cls].rc;
            const bool is_g3 = grp.kind == PANTEA_G3;
            for (int e = lane; e < count; e += 32) {
                const int jk = lists[lo + e];
                const int nj = jk & 0xffff, nk = jk >> 16;
                const T* pj = snb + (size_t)nj * stride;
                const T* pk = snb + (size_t)nk * stride;
                const T uxj = pj[0], uyj = pj[1], uzj = pj[2], rj = pj[3], ivj = pj[4], fcj = pj[fco], dfj = pj[fco + 1];
                const T uxk = pk[0], uyk = pk[1], uzk = pk[2], rk = pk[3], ivk = pk[4], fck = pk[fco], dfk = pk[fco + 1];
                T ex = uxj * rj - uxk * rk, ey = uyj * rj - uyk * rk, ez = uzj * rj - uzk * rk;  // d_jk = d_ij - d_ik
                if (wrap_jk) { ex = min_image(ex, lx); ey = min_image(ey, ly); ez = min_image(ez, lz); }
                const T rjk2 = ex * ex + ey * ey + ez * ez;
                if (!(rjk2 > (T)0)) continue;  // k == j / coincident neighbours excluded (acsf.py:325)
                T fcjk = (T)1, dfjk = (T)0, rjk = (T)0, ivjk = (T)0, r2 = rj * rj + rk * rk;
                if (is_g3) {
                    rjk = t_sqrt<T>(rjk2); ivjk = (T)1 / rjk;
                    cutoff_eval_ool<T>(ctype, rjk, rc, &fcjk, &dfjk);
                    r2 += rjk2;
                    if (fcjk == (T)0 && dfjk == (T)0) continue;
                }
                const T cost = uxj * uxk + uyj * uyk + uzj * uzk;
                T a1 = 0, a2 = 0, a3 = 0, b1 = 0, b2 = 0;  // g_j = a1 u_j + a2 u_k + a3 u_jk ; g_k = b1 u_k + b2 u_j - a3 u_jk
                for (int m = 0; m < grp.count; ++m) {
                    const AngularMember mem = tab.members[grp.first + m];
                    const T wm = sw[mem.out];
                    const T eta = (T)mem.eta, lam = (T)mem.lambda0;
                    const T bs = (T)1 + lam * cost;
                    T pw1 = (T)1;
                    if (mem.izeta != 1) pw1 = mem.izeta > 1 ? powi<T>(bs, mem.izeta - 1) : pow_general<T>(bs, (T)(mem.zeta - 1.0));
                    const bool z0 = mem.izeta == 0;  // x^0 = 1 also at x = 0 (the reference's pow)
                    const T ee = z0 ? (T)0 : t_exp<T>(-eta * r2) * pw1 * (T)mem.pref;  // pref (1 + lambda c)^(zeta-1) exp(-eta r2)
                    const T A = z0 ? t_exp<T>(-eta * r2) * (T)mem.pref : ee * bs;      // pref (1 + lambda c)^zeta exp(-eta r2)
                    const T Tc = wm * (T)(mem.zeta * mem.lambda0) * ee * (fcj * fck * fcjk);
                    const T Tj = wm * A * fck * fcjk * (dfj - (T)2 * eta * rj * fcj);
                    const T Tk = wm * A * fcj * fcjk * (dfk - (T)2 * eta * rk * fck);
                    a2 += Tc * ivj; a1 += Tj - Tc * cost * ivj;
                    b2 += Tc * ivk; b1 += Tk - Tc * cost * ivk;
                    if (is_g3) a3 += wm * A * fcj * fck * (dfjk - (T)2 * eta * rjk * fcjk);
                }
                const T ujx = ex * ivjk, ujy = ey * ivjk, ujz = ez * ivjk;  // u_jk (zero for G9)
                atomicAdd(&sacc[3 * nj], a1 * uxj + a2 * uxk + a3 * ujx);
                atomicAdd(&sacc[3 * nj + 1], a1 * uyj + a2 * uyk + a3 * ujy);
                atomicAdd(&sacc[3 * nj + 2], a1 * uzj + a2 * uzk + a3 * ujz);
                atomicAdd(&sacc[3 * nk], b1 * uxk + b2 * uxj - a3 * ujx);
                atomicAdd(&sacc[3 * nk + 1], b1 * uyk + b2 * uyj - a3 * ujy);
                atomicAdd(&sacc[3 * nk + 2], b1 * uzk + b2 * uzj - a3 * ujz);
            }
            __syncwarp();
        }
    }

    // ---- scatter: F_n += g_in, F_i -= sum_n g_in --------------------------------------------------------
    T cx = 0, cy = 0, cz = 0;
    for (int n = lane; n < total; n += 32) {
        const T gx = sacc[3 * n], gy = sacc[3 * n + 1], gz = sacc[3 * n + 2];
        T* f = a.forces + (size_t)3 * sidx[n];
        atomicAdd(f, gx); atomicAdd(f + 1, gy); atomicAdd(f + 2, gz);
        cx += gx; cy += gy; cz += gz;
    }
    cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz);
    if (lane == 0) {
        T* f = a.forces + (size_t)3 * out_row;
        atomicAdd(f, -cx); atomicAdd(f + 1, -cy); atomicAdd(f + 2, -cz);
    }
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
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: `configured` is indexed by the current device
int opt_in_smem(const void* kern, size_t smem, size_t* configured, const char* what) {
    int dev = 0;
    PANTEA_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(PANTEA_EINVAL, "device index out of range");
    if (smem > configured[dev]) {
        if (smem > 227 * 1024) return fail(PANTEA_EINVAL, what);
        PANTEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    return PANTEA_OK;
}
static int g_num_sms_dev[64] = {0};  // per device
// pantea_set_fast_path: 1 (default; PANTEA_EVAL_V2=0 in the environment starts with 0) lets qualifying evaluations take
// the specialised kernels of acsf2.cu, 0 keeps everything on the generic kernels of this file
// pantea_set_gauss_screen: threshold T of the fast path's Gaussian screening (0: off)
static std::atomic<double> g_gauss_screen{[]() { const char* e = std::getenv("PANTEA_GAUSS_SCREEN"); return e ? std::atof(e) : 40.0; }()};
static std::atomic<int> g_fast_path{[]() { const char* e = std::getenv("PANTEA_EVAL_V2"); return (e && e[0] == '0') ? 0 : 1; }()};

template <typename T, int WPA, bool GRAD, int MCH>
static int launch_eval(const AtomArgs<T>& args, cudaStream_t st) {
    const int apb = kEvalWarps / WPA;
    const size_t smem = apb * eval_smem_bytes<T>(args.scap, args.n_cls_max, args.n_sf_max, args.n_neurons_max, args.width_max, WPA);
    auto kern = hdnnp_eval_kernel<T, WPA, GRAD, MCH>;
    static size_t configured[64] = {0};  // per instantiation and device
    int rc_s = opt_in_smem((const void*)kern, smem, configured, "neighbour capacity / potential too large for shared memory");
    if (rc_s != PANTEA_OK) return rc_s;
    const int blocks = (args.n_work + apb - 1) / apb;
    kern<<<blocks, apb * WPA * 32, smem, st>>>(args);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

template <typename T, int WPA, bool GRAD>
static int launch_mch(const AtomArgs<T>& a, int max_members, cudaStream_t st) {
    if (max_members <= 1) return launch_eval<T, WPA, GRAD, 1>(a, st);
    return launch_eval<T, WPA, GRAD, 4>(a, st);
}

__global__ void skin_lists_fresh_kernel(int32_t* __restrict__ skin_flags) { skin_flags[1] = 0; }

template <typename T>
static int launch_filter(const AtomArgs<T>& a, cudaStream_t st) {
    const size_t smem = (size_t)kFilterWarps * (a.scap + 1) * sizeof(float4);
    auto kern = pair_filter_kernel<T>;
    static size_t configured[64] = {0};
    int rc_s = opt_in_smem((const void*)kern, smem, configured, "neighbour capacity too large for shared memory");
    if (rc_s != PANTEA_OK) return rc_s;
    const int blocks = (a.n_work + kFilterWarps - 1) / kFilterWarps;
    kern<<<blocks, kFilterWarps * 32, smem, st>>>(a);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

static int launch_v2_dispatch(const AtomArgs<double>& a, cudaStream_t st) { return launch_v2(a, st); }
static int launch_v2_dispatch(const AtomArgs<float>&, cudaStream_t) { return fail(PANTEA_EINVAL, "fast path is double precision only"); }

// pair-list storage: [max_atoms][pair_cap] entries + [max_atoms][max_groups + 1] offsets
static int ensure_pair_storage(pantea_workspace* ws) {
    const pantea_potential* pot = ws->pot;
    const int scap = ws->smem_cap > 0 && ws->smem_cap < ws->cap ? ws->smem_cap : ws->cap;
    int want = ws->pair_cap_request > 0 ? ws->pair_cap_request : 24 * scap;  // first guess; refined from observed maxima
    want = (want + 31) / 32 * 32;
    if (ws->pairs && want == ws->pair_cap && ws->pair_groups == pot->max_groups) return PANTEA_OK;
    if (ws->pairs) cudaFree(ws->pairs);
    if (ws->pair_off) cudaFree(ws->pair_off);
    ws->pairs = ws->pair_off = nullptr;
    cudaError_t err = cudaMalloc((void**)&ws->pairs, sizeof(int32_t) * ((size_t)ws->max_atoms * want + 1024));
    if (err == cudaSuccess) err = cudaMalloc((void**)&ws->pair_off, sizeof(int32_t) * (size_t)ws->max_atoms * (pot->max_groups + 1));
    if (err != cudaSuccess)
        return fail(err == cudaErrorMemoryAllocation ? PANTEA_ENOMEM : PANTEA_ECUDA,
                    std::string("pair-list allocation: ") + cudaGetErrorString(err));
    ws->pair_cap = want;
    ws->pair_groups = pot->max_groups;
    ws->lists_valid = false;  // (Verlet skin) fresh storage: the next energy pass filters again
    ++ws->arg_epoch;
    return PANTEA_OK;
}

template <typename T>
static int atom_kernel_typed(pantea_workspace* ws, int element_slot, const int32_t* centres, int64_t n_centres, void* G,
                             void* dG, void* e_atom, void* forces_out, int force_mode, cudaStream_t st) {
    // full-force mode: the evaluation only needs values, the network kernel hands out dE_i/dG_is, a scatter pass follows
    const bool full = forces_out != nullptr && force_mode == PANTEA_FORCE_FULL;
    void* forces = full ? nullptr : forces_out;
    const pantea_potential* pot = ws->pot;
    int rc = ensure_pair_storage(ws);
    if (rc != PANTEA_OK) return rc;
    AtomArgs<T> a{};
    a.rec = (const Rec<T>*)ws->rec; a.nbr = ws->nbr; a.tcount = ws->nbr_tcount; a.cap = ws->cap;
    a.scap = ws->smem_cap > 0 && ws->smem_cap < ws->cap ? ws->smem_cap : ws->cap; a.flags = ws->flags;
    a.slot_of = ws->slot_of; a.struct_of = ws->struct_of; a.boxes = ws->boxes;
    a.box = BoxArgK{ws->box[0], ws->box[1], ws->box[2], ws->has_box ? 1 : 0};
    double lmin = ws->box[0] < ws->box[1] ? ws->box[0] : ws->box[1];
    if (ws->box[2] < lmin) lmin = ws->box[2];
    a.wrap_jk = ws->boxes ? 1 : (ws->has_box && 0.5 * lmin < 2.0 * ws->rc * (1.0 + 1e-9) ? 1 : 0);
    a.rc_list = ws->rc;
    a.skin = ws->skin_active ? (float)ws->skin : 0.f;
    a.rec_bytes = ws->compute32 ? 48 : 80;
    a.dup_flag = ws->mode == kModeCell ? ws->wide_flag + 1 : nullptr;
    a.dup_always = ws->mode == kModeCell ? 0 : 1;
    a.screen_t = a.skin > 0.f ? 0.f : (float)g_gauss_screen.load(std::memory_order_relaxed);  // (kept lists must not depend on the positions)
    a.filter_guard = nullptr;
    a.tables = pot->dev; a.n_types = pot->n_elements; a.element_slot = element_slot;
    a.centres = centres;
    const bool energy_pass = (e_atom || forces_out) && !G && !dG;
    a.by_slot = energy_pass ? 1 : 0;
    a.own_begin = (int)ws->own_begin; a.own_end = ws->own_end < 0 ? (int)ws->n : (int)ws->own_end;
    a.owned_slots = energy_pass && ws->owned_active ? ws->owned_slots : nullptr;
    a.n_work = energy_pass ? (a.owned_slots ? a.own_end - a.own_begin : (int)ws->n) : (int)n_centres;
    a.n_work_dev = nullptr;
    if (ws->role) {  // brick decomposition: owned atoms by role, their number on the device
        if (!energy_pass || !ws->owned_active) return fail(PANTEA_EINVAL, "role-owned workspaces only run the energy / force pass");
        a.n_work = (int)ws->own_cap;
        a.n_work_dev = ws->cell_own + (int64_t)ws->ncell[0] * ws->ncell[1] * ws->ncell[2];
    }
    a.pairs = ws->pairs; a.pair_off = ws->pair_off; a.pair_cap = ws->pair_cap; a.max_groups = pot->max_groups;
    a.G = (T*)G; a.dG = (T*)dG; a.e_atom = (T*)e_atom; a.forces = (T*)forces;
    a.gbuf = nullptr;
    a.wbuf = nullptr;
    if (full) {
        if (!ws->wbuf) {
            const size_t bytes = sizeof(T) * (size_t)ws->max_atoms * (pot->max_sf > 0 ? pot->max_sf : 1);
            cudaError_t err = cudaMalloc(&ws->wbuf, bytes);
            ++ws->arg_epoch;
            if (err != cudaSuccess) return fail(PANTEA_ENOMEM, std::string("network-gradient buffer: ") + cudaGetErrorString(err));
        }
        a.wbuf = (T*)ws->wbuf;
    }
    if (e_atom || forces_out) {
        if (!ws->gbuf) {
            const size_t bytes = sizeof(T) * (size_t)ws->max_atoms * (pot->max_sf > 0 ? pot->max_sf : 1) * 4;
            cudaError_t err = cudaMalloc(&ws->gbuf, bytes);
            ++ws->arg_epoch;
            if (err != cudaSuccess) return fail(PANTEA_ENOMEM, std::string("descriptor hand-off buffer: ") + cudaGetErrorString(err));
        }
        a.gbuf = (T*)ws->gbuf;
    }
    a.counters = ws->counters;
    a.g_stride = element_slot >= 0 ? pot->host[element_slot].n_sf : pot->max_sf;
    a.n_cls_max = pot->max_cls; a.n_sf_max = pot->max_sf > 0 ? pot->max_sf : 1;
    a.n_neurons_max = pot->max_neurons; a.width_max = pot->max_width;
    if (a.n_work == 0) return PANTEA_OK;
    if (a.n_work > ws->max_atoms) return fail(PANTEA_EINVAL, "more centres than the workspace capacity");
    int dev_now = 0;
    PANTEA_CUDA_TRY(cudaGetDevice(&dev_now));
    if (dev_now < 0 || dev_now >= 64) return fail(PANTEA_EINVAL, "device index out of range");
    if (g_num_sms_dev[dev_now] == 0)
        PANTEA_CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms_dev[dev_now], cudaDevAttrMultiProcessorCount, dev_now));
    const int g_num_sms = g_num_sms_dev[dev_now];
    const bool grad = dG != nullptr || forces != nullptr;
    // few atoms: several warps per atom so that every SM sub-partition has work.  The energy pass decides on the
    // system size, not on this rank's share: the summation order -- hence every bit of the result -- is then the same
    // on 1 and on N GPUs
    const bool wide = (int64_t)(energy_pass ? (int)ws->n : a.n_work) < (int64_t)g_num_sms * 64;
    // fast path (acsf2.cu): double precision with gradients, cell-list rows of a box that needs no minimum image on
    // r_jk, a potential whose tables qualify, no work counters; its pair lists have their own format
    const bool use_v2 = g_fast_path.load(std::memory_order_relaxed) != 0 && sizeof(T) == 8 && pot->v2_ok && grad && !wide && !a.wrap_jk &&
                        ws->mode == kModeCell && (a.scap + 1) * 80 < 65536;
    if (ws->lists_v2 != use_v2) ws->lists_valid = false;  // (Verlet skin) lists kept from a pass in the other format
    ws->lists_v2 = use_v2;
    if (pot->max_groups > 0 || use_v2) {
        // Verlet skin: the energy pass keeps its pair lists until the neighbour rows are rebuilt (device flag)
        const bool reuse = energy_pass && ws->skin_active;
        if (reuse) {
            if (!ws->lists_valid) PANTEA_CUDA_TRY(cudaMemsetAsync(ws->skin_flags + 1, 1, 4, st));
            a.filter_guard = ws->skin_flags + 1;
        }
        if (use_v2) rc = launch_v2_dispatch(a, st);
        else rc = launch_filter<T>(a, st);
        if (rc != PANTEA_OK) return rc;
        if (reuse) {
            skin_lists_fresh_kernel<<<1, 1, 0, st>>>(ws->skin_flags);
            PANTEA_LAUNCH_CHECK();
        }
        ws->lists_valid = reuse;
    }
    const int mm = pot->max_members;
    if (use_v2) rc = PANTEA_OK;  // filter and evaluation were launched together above
    else if (wide) rc = grad ? launch_mch<T, 4, true>(a, mm, st) : launch_mch<T, 4, false>(a, mm, st);
    else rc = grad ? launch_mch<T, PANTEA_EVAL_WPA, true>(a, mm, st) : launch_mch<T, PANTEA_EVAL_WPA, false>(a, mm, st);
    if (rc != PANTEA_OK || !a.gbuf) return rc;
    {
        const int gw = a.width_max > a.n_sf_max ? a.width_max : a.n_sf_max;
        const size_t smem = (size_t)(a.n_sf_max + 2 * a.n_neurons_max + 2 * gw) * kMlpThreads * sizeof(T);
        auto kern = mlp_force_kernel<T>;
        static size_t configured[64] = {0};
        int rc_s = opt_in_smem((const void*)kern, smem, configured, "network too wide for the per-thread shared-memory scratch");
        if (rc_s != PANTEA_OK) return rc_s;
        kern<<<(a.n_work + kMlpThreads - 1) / kMlpThreads, kMlpThreads, smem, st>>>(a);
        PANTEA_LAUNCH_CHECK();
    }
    if (full) {
        PANTEA_CUDA_TRY(cudaMemsetAsync(forces_out, 0, sizeof(T) * 3 * (size_t)ws->n, st));
        a.forces = (T*)forces_out;
        const size_t smem = (size_t)kFullWarps * full_smem_bytes<T>(a.scap, a.n_cls_max, a.n_sf_max);
        auto kern = full_force_kernel<T>;
        static size_t configured[64] = {0};
        int rc_s = opt_in_smem((const void*)kern, smem, configured, "neighbour capacity too large for the full-force pass");
        if (rc_s != PANTEA_OK) return rc_s;
        kern<<<(a.n_work + kFullWarps - 1) / kFullWarps, kFullWarps * 32, smem, st>>>(a);
        PANTEA_LAUNCH_CHECK();
    }
    return PANTEA_OK;
}

int atom_kernel_launch(pantea_workspace* ws, int element_slot, const int32_t* centres, int64_t n_centres, void* G,
                       void* dG, void* e_atom, void* forces, cudaStream_t st, int force_mode) {
    if (ws->cap > 0xffff) return fail(PANTEA_EINVAL, "max_neighbors must be < 65536");
    if (ws->dtype == PANTEA_F64) return atom_kernel_typed<double>(ws, element_slot, centres, n_centres, G, dG, e_atom, forces, force_mode, st);
    return atom_kernel_typed<float>(ws, element_slot, centres, n_centres, G, dG, e_atom, forces, force_mode, st);
}

}  // namespace pantea

using namespace pantea;

extern "C" {

int pantea_workspace_set_compute_precision(pantea_workspace* ws, int32_t bits) {
    if (!ws) return fail(PANTEA_EINVAL, "pantea_workspace_set_compute_precision: NULL workspace");
    if (bits != 32 && bits != 64) return fail(PANTEA_EINVAL, "pantea_workspace_set_compute_precision: bits must be 32 or 64");
    if (bits == 32 && ws->dtype != PANTEA_F64) return fail(PANTEA_EINVAL, "pantea_workspace_set_compute_precision: the mixed mode belongs to PANTEA_F64 workspaces (a PANTEA_F32 workspace computes in single precision throughout)");
    const bool want = bits == 32;
    if (want != ws->compute32) { ws->compute32 = want; ws->lists_valid = false; ++ws->arg_epoch; }
    return PANTEA_OK;
}

int pantea_set_fast_path(int32_t enable) {
    const int old = g_fast_path.exchange(enable ? 1 : 0);
    return old;
}

double pantea_set_gauss_screen(double threshold) {
    return g_gauss_screen.exchange(threshold > 0.0 ? threshold : 0.0);
}

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
    if (force_mode != PANTEA_FORCE_REFERENCE && force_mode != PANTEA_FORCE_FULL)
        return fail(PANTEA_EINVAL, "pantea_energy_forces: unknown force_mode");
    if (!e_atom && !forces && !e_total) return fail(PANTEA_EINVAL, "pantea_energy_forces: all outputs are NULL");
    if (ws->n == 0) return PANTEA_OK;
    void* ea = e_atom ? e_atom : (e_total ? ws->md_eatom : nullptr);
    int rc = atom_kernel_launch(ws, -1, nullptr, 0, nullptr, nullptr, ea, forces, (cudaStream_t)stream, force_mode);
    if (rc != PANTEA_OK) return rc;
    if (e_total) return reduce_energy(ws, ea, e_total, (cudaStream_t)stream);
    return PANTEA_OK;
}

}  // extern "C"
