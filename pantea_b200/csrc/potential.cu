// Potential tables, workspace lifetime and error plumbing of libpantea_b200.so.
#include <atomic>
#include <cmath>
#include <cstring>

#include "internal.cuh"

namespace pantea {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

static bool valid_cutoff(int t) { return t >= PANTEA_CUT_HARD && t <= PANTEA_CUT_POLY2; }
static bool valid_act(int a) { return a >= PANTEA_ACT_IDENTITY && a <= PANTEA_ACT_HARMONIC; }

// Translate one host element descriptor into the device table layout:
//   radial functions keep their order; angular functions are grouped by
//   (type_j, type_k, cutoff class, kind) so that one triplet evaluation serves every member.
static int build_table(const pantea_element_desc& d, int n_elements, int e, ElementTable& t, std::string& why) {
    std::memset(&t, 0, sizeof(t));
    if (d.n_symfunc < 0 || d.n_symfunc > kMaxSF) { why = "n_symfunc out of range"; return PANTEA_EINVAL; }
    if (d.n_symfunc > 0 && !d.symfunc) { why = "symfunc is NULL"; return PANTEA_EINVAL; }
    t.n_sf = d.n_symfunc;
    auto find_cls = [&](int type, double rc) -> int {
        for (int c = 0; c < t.n_cls; ++c)
            if (t.cls[c].type == type && t.cls[c].rc == rc) return c;
        if (t.n_cls == kMaxCut) return -1;
        t.cls[t.n_cls] = {type, rc};
        return t.n_cls++;
    };
    bool seen_angular = false;
    int n_members = 0;
    for (int s = 0; s < d.n_symfunc; ++s) {
        const pantea_symfunc_desc& sf = d.symfunc[s];
        if (!valid_cutoff(sf.cutoff_type)) { why = "unknown cutoff type"; return PANTEA_EINVAL; }
        if (!(sf.r_cutoff > 0.0)) { why = "r_cutoff must be positive"; return PANTEA_EINVAL; }
        if (sf.type_j < 1 || sf.type_j > n_elements) { why = "type_j outside 1..n_elements"; return PANTEA_EINVAL; }
        int cls = find_cls(sf.cutoff_type, sf.r_cutoff);
        if (cls < 0) { why = "more than PANTEA_MAX_CUTOFFS distinct cutoffs in one element"; return PANTEA_EINVAL; }
        if (sf.r_cutoff > t.rc_max) t.rc_max = sf.r_cutoff;
        if (sf.kind == PANTEA_G1 || sf.kind == PANTEA_G2) {
            if (seen_angular) { why = "radial symmetry functions must precede angular ones"; return PANTEA_EINVAL; }
            RadialSF& r = t.radial[t.n_radial++];
            r.type_j = sf.type_j - 1; r.kind = sf.kind; r.cls = cls; r.out = s; r.eta = sf.eta; r.r_shift = sf.r_shift;
        } else if (sf.kind == PANTEA_G3 || sf.kind == PANTEA_G9) {
            seen_angular = true;
            if (sf.type_k < 1 || sf.type_k > n_elements) { why = "type_k outside 1..n_elements"; return PANTEA_EINVAL; }
            ++n_members;
        } else { why = "unknown symmetry function kind"; return PANTEA_EINVAL; }
    }
    // group angular members (stable in declaration order)
    std::vector<char> used(d.n_symfunc, 0);
    int m = 0;
    for (int s = t.n_radial; s < d.n_symfunc; ++s) {
        if (used[s]) continue;
        const pantea_symfunc_desc& a = d.symfunc[s];
        int cls = find_cls(a.cutoff_type, a.r_cutoff);
        if (t.n_groups == kMaxGroups) { why = "too many angular groups"; return PANTEA_EINVAL; }
        AngularGroup& g = t.groups[t.n_groups++];
        g.type_j = a.type_j - 1; g.type_k = a.type_k - 1; g.cls = cls; g.kind = a.kind; g.first = m; g.count = 0;
        for (int q = s; q < d.n_symfunc; ++q) {
            const pantea_symfunc_desc& b = d.symfunc[q];
            if (used[q] || b.kind != a.kind || b.type_j != a.type_j || b.type_k != a.type_k ||
                b.cutoff_type != a.cutoff_type || b.r_cutoff != a.r_cutoff) continue;
            used[q] = 1;
            AngularMember& mem = t.members[m++];
            mem.eta = b.eta; mem.lambda0 = b.lambda0; mem.zeta = b.zeta; mem.pref = std::pow(2.0, 1.0 - b.zeta);
            double iz = std::floor(b.zeta);
            mem.izeta = (iz == b.zeta && b.zeta >= 0.0 && b.zeta <= 64.0) ? (int)iz : -1;
            mem.out = q;
            ++g.count;
        }
    }
    (void)n_members; (void)e;
    // fast-path eligibility (acsf2.cu)
    t.v2_ok = 1;
    for (int b = 0; b < kBuckets; ++b) { t.v2_group_of_type[b] = -1; t.v2_eta[b] = 0.0; t.v2_wscale[b] = 0.0; }
    if (t.n_cls > 1) t.v2_ok = 0;
    if (t.n_cls == 1 && t.cls[0].type != PANTEA_CUT_TANHU && t.n_groups > 0) t.v2_ok = 0;
    for (int g = 0; g < t.n_groups && t.v2_ok; ++g) {
        const AngularGroup& grp = t.groups[g];
        const AngularMember& mem = t.members[grp.first];
        if (grp.count != 1 || grp.kind != PANTEA_G3 || mem.izeta < 1 || mem.izeta > 16 || !(mem.eta >= 0.0) ||
            !(mem.eta * t.cls[grp.cls].rc * t.cls[grp.cls].rc * 1.01 < 690.0)) { t.v2_ok = 0; break; }
        for (int side = 0; side < 2; ++side) {
            const int ty = side == 0 ? grp.type_j : grp.type_k;
            if (side == 1 && grp.type_k == grp.type_j) break;
            if (t.v2_group_of_type[ty] >= 0) t.v2_ok = 0;  // a neighbour type in two groups needs two weight pairs
            t.v2_group_of_type[ty] = g;
            t.v2_eta[ty] = mem.eta;
            t.v2_wscale[ty] = std::sqrt(mem.pref);
        }
    }
    // scaler
    t.has_scaler = (d.scale_shift && d.scale_slope && d.scale_offset) ? 1 : 0;
    for (int s = 0; s < d.n_symfunc; ++s) {
        t.shift[s] = t.has_scaler ? d.scale_shift[s] : 0.0;
        t.slope[s] = t.has_scaler ? d.scale_slope[s] : 1.0;
        t.offset[s] = t.has_scaler ? d.scale_offset[s] : 0.0;
    }
    // network
    if (d.n_layers < 0 || d.n_layers > kMaxLayers) { why = "n_layers out of range"; return PANTEA_EINVAL; }
    t.n_layers = d.n_layers;
    if (d.n_layers > 0) {
        if (!d.layer_sizes || !d.activations || !d.weights) { why = "network arrays are NULL"; return PANTEA_EINVAL; }
        if (d.layer_sizes[0] != d.n_symfunc) { why = "layer_sizes[0] must equal n_symfunc"; return PANTEA_EINVAL; }
        if (d.layer_sizes[d.n_layers] != 1) { why = "the output layer must have one neuron"; return PANTEA_EINVAL; }
        int off = 0;
        for (int l = 0; l <= d.n_layers; ++l) {
            if (d.layer_sizes[l] < 1 || d.layer_sizes[l] > 1024) { why = "layer size out of range"; return PANTEA_EINVAL; }
            t.sizes[l] = d.layer_sizes[l];
        }
        for (int l = 0; l < d.n_layers; ++l) {
            if (!valid_act(d.activations[l])) { why = "unknown activation"; return PANTEA_EINVAL; }
            t.acts[l] = d.activations[l];
            t.w_off[l] = off;
            off += t.sizes[l] * t.sizes[l + 1] + t.sizes[l + 1];
            t.n_neurons += t.sizes[l + 1];
        }
    }
    return PANTEA_OK;
}

static int64_t weight_count(const ElementTable& t) {
    int64_t n = 0;
    for (int l = 0; l < t.n_layers; ++l) n += (int64_t)t.sizes[l] * t.sizes[l + 1] + t.sizes[l + 1];
    return n;
}

}  // namespace pantea

using namespace pantea;

extern "C" {

const char* pantea_last_error(void) { return g_last_error.c_str(); }
const char* pantea_version(void) { return "pantea_b200 0.1.0 (sm_100a)"; }
int64_t pantea_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int pantea_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pantea_potential_create(const pantea_potential_desc* desc, pantea_potential** out) {
    if (!desc || !out) return fail(PANTEA_EINVAL, "pantea_potential_create: NULL argument");
    if (desc->n_elements < 1 || desc->n_elements > kMaxTypes)
        return fail(PANTEA_EINVAL, "pantea_potential_create: n_elements must be in 1..PANTEA_MAX_TYPES");
    if (pantea_device_count() < 1) return fail(PANTEA_ECUDA, "pantea_potential_create: no CUDA device (no CPU fallback)");
    auto* pot = new pantea_potential();
    pot->n_elements = desc->n_elements;
    pot->host.resize(desc->n_elements);
    pot->dev_weights.assign(desc->n_elements, nullptr);
    for (int e = 0; e < desc->n_elements; ++e) {
        std::string why;
        int rc = build_table(desc->elements[e], desc->n_elements, e, pot->host[e], why);
        if (rc != PANTEA_OK) {
            pantea_potential_destroy(pot);
            return fail(rc, "pantea_potential_create: element " + std::to_string(e) + ": " + why);
        }
        ElementTable& t = pot->host[e];
        if (t.rc_max > pot->rc_max) pot->rc_max = t.rc_max;
        if (t.n_sf > pot->max_sf) pot->max_sf = t.n_sf;
        if (t.n_cls > pot->max_cls) pot->max_cls = t.n_cls;
        for (int g = 0; g < t.n_groups; ++g) if (t.groups[g].count > pot->max_members) pot->max_members = t.groups[g].count;
        if (t.n_groups > pot->max_groups) pot->max_groups = t.n_groups;
        if (!t.v2_ok) pot->v2_ok = false;
        if (t.n_neurons > pot->max_neurons) pot->max_neurons = t.n_neurons;
        for (int l = 0; l <= t.n_layers; ++l)
            if (t.n_layers > 0 && t.sizes[l] > pot->max_width) pot->max_width = t.sizes[l];
        int64_t nw = weight_count(t);
        if (nw > 0) {
            cudaError_t err = cudaMalloc(&pot->dev_weights[e], sizeof(double) * nw);
            if (err == cudaSuccess)
                err = cudaMemcpy(pot->dev_weights[e], desc->elements[e].weights, sizeof(double) * nw, cudaMemcpyHostToDevice);
            if (err != cudaSuccess) {
                pantea_potential_destroy(pot);
                return fail(PANTEA_ECUDA, std::string("pantea_potential_create: ") + cudaGetErrorString(err));
            }
            t.weights = pot->dev_weights[e];
        }
    }
    cudaError_t err = cudaMalloc(&pot->dev, sizeof(ElementTable) * pot->n_elements);
    if (err == cudaSuccess)
        err = cudaMemcpy(pot->dev, pot->host.data(), sizeof(ElementTable) * pot->n_elements, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        pantea_potential_destroy(pot);
        return fail(PANTEA_ECUDA, std::string("pantea_potential_create: ") + cudaGetErrorString(err));
    }
    *out = pot;
    return PANTEA_OK;
}

int pantea_potential_destroy(pantea_potential* pot) {
    if (!pot) return PANTEA_OK;
    for (double* w : pot->dev_weights)
        if (w) cudaFree(w);
    if (pot->dev) cudaFree(pot->dev);
    delete pot;
    return PANTEA_OK;
}

double pantea_potential_cutoff(const pantea_potential* pot) { return pot ? pot->rc_max : 0.0; }

int pantea_workspace_create(const pantea_potential* pot, int64_t max_atoms, int32_t max_neighbors, int32_t dtype,
                            pantea_workspace** out) {
    if (!out) return fail(PANTEA_EINVAL, "pantea_workspace_create: NULL out");
    if (dtype != PANTEA_F64 && dtype != PANTEA_F32) return fail(PANTEA_EINVAL, "pantea_workspace_create: dtype must be 64 or 32");
    if (max_atoms < 1 || max_atoms >= (int64_t(1) << 28)) return fail(PANTEA_EINVAL, "pantea_workspace_create: max_atoms out of range");
    if (max_neighbors < 1 || max_neighbors > 4096) return fail(PANTEA_EINVAL, "pantea_workspace_create: max_neighbors out of range");
    if (pantea_device_count() < 1) return fail(PANTEA_ECUDA, "pantea_workspace_create: no CUDA device (no CPU fallback)");
    auto* ws = new pantea_workspace();
    ws->pot = pot;
    ws->max_atoms = max_atoms;
    ws->cap = (max_neighbors + 31) / 32 * 32;
    ws->dtype = dtype;
    ws->n_types = pot ? pot->n_elements : 0;
    const size_t esz = dtype == PANTEA_F64 ? 8 : 4;
    const size_t rsz = dtype == PANTEA_F64 ? sizeof(Rec<double>) : sizeof(Rec<float>);
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) {
        if (err == cudaSuccess) err = cudaMalloc(p, bytes);
    };
    alloc(&ws->rec, rsz * max_atoms);
    if (dtype == PANTEA_F64) alloc(&ws->rec_screen, sizeof(Rec<float>) * max_atoms);
    alloc((void**)&ws->wide_flag, 8);  // [0] atoms far outside the box, [1] coincident atoms present
    alloc(&ws->pos_ref, esz * 3 * max_atoms);
    alloc((void**)&ws->skin_flags, 4 * 4);
    alloc((void**)&ws->slot_of, 4 * max_atoms);
    alloc((void**)&ws->struct_of, 4 * max_atoms);
    alloc((void**)&ws->nbr, 4 * (size_t)max_atoms * ws->cap);
    alloc((void**)&ws->nbr_tcount, 4 * (size_t)max_atoms * kBuckets);
    alloc((void**)&ws->cell_of, 4 * max_atoms);
    alloc((void**)&ws->tmp_order, 4 * max_atoms);
    alloc((void**)&ws->owned_slots, 4 * max_atoms);
    alloc((void**)&ws->flags, 4 * 4);
    ws->e_partial_cap = 4096;
    alloc((void**)&ws->e_partial, 8 * ws->e_partial_cap);
    alloc(&ws->md_forces, esz * 3 * max_atoms);
    alloc(&ws->md_eatom, esz * max_atoms);
    alloc((void**)&ws->md_ke, 8 * 4);
    if (err == cudaSuccess) err = cudaMemset(ws->flags, 0, 16);
    if (err == cudaSuccess) err = cudaMemset(ws->skin_flags, 0, 16);
    if (err != cudaSuccess) {
        pantea_workspace_destroy(ws);
        return fail(err == cudaErrorMemoryAllocation ? PANTEA_ENOMEM : PANTEA_ECUDA,
                    std::string("pantea_workspace_create: ") + cudaGetErrorString(err));
    }
    *out = ws;
    return PANTEA_OK;
}

int pantea_workspace_destroy(pantea_workspace* ws) {
    if (!ws) return PANTEA_OK;
    if (ws->md_graph) cudaGraphExecDestroy(ws->md_graph);
    if (ws->capture_stream) cudaStreamDestroy(ws->capture_stream);
    void* ptrs[] = {ws->rec, ws->slot_of, ws->struct_of, ws->nbr, ws->nbr_tcount, ws->cell_of, ws->tmp_order,
                    ws->cell_start, ws->cell_fill, ws->flags, ws->e_partial, ws->md_forces, ws->md_eatom, ws->md_ke,
                    ws->pairs, ws->pair_off, ws->gbuf, ws->wbuf, ws->scan_sums, ws->cell_own_cnt, ws->rec_screen, ws->wide_flag, ws->pos_ref, ws->skin_flags, ws->owned_slots, ws->cell_own};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete ws;
    return PANTEA_OK;
}

int pantea_workspace_set_skin(pantea_workspace* ws, double skin) {
    if (!ws) return fail(PANTEA_EINVAL, "pantea_workspace_set_skin: NULL workspace");
    if (!(skin >= 0.0)) return fail(PANTEA_EINVAL, "pantea_workspace_set_skin: skin must be >= 0");
    if (skin != ws->skin) {
        ws->skin = skin;
        ++ws->arg_epoch;
        ws->skin_active = false;  // the next build gathers fresh rows
        ws->lists_valid = false;
    }
    return PANTEA_OK;
}

int pantea_neighbor_rebuilds(pantea_workspace* ws, int64_t* builds, void* stream) {
    if (!ws || !builds) return fail(PANTEA_EINVAL, "pantea_neighbor_rebuilds: NULL argument");
    int32_t h[4] = {0, 0, 0, 0};
    PANTEA_CUDA_TRY(cudaMemcpyAsync(h, ws->skin_flags, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PANTEA_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    builds[0] = h[3];
    builds[1] = h[2];
    return PANTEA_OK;
}

int pantea_workspace_set_owned_range(pantea_workspace* ws, int64_t begin, int64_t end) {
    if (!ws) return fail(PANTEA_EINVAL, "pantea_workspace_set_owned_range: NULL workspace");
    if (end >= 0 && (begin < 0 || begin > end)) return fail(PANTEA_EINVAL, "pantea_workspace_set_owned_range: bad range");
    if (ws->own_begin != begin || ws->own_end != end) ++ws->arg_epoch;
    ws->own_begin = begin;
    ws->own_end = end;
    return PANTEA_OK;
}

}  // extern "C"

namespace pantea {
int ensure_cell_capacity(pantea_workspace* ws, int64_t ncells) {
    if (ncells <= ws->cell_cap) return PANTEA_OK;
    if (ws->cell_start) cudaFree(ws->cell_start);
    if (ws->cell_fill) cudaFree(ws->cell_fill);
    if (ws->cell_own) cudaFree(ws->cell_own);
    if (ws->scan_sums) cudaFree(ws->scan_sums);
    if (ws->cell_own_cnt) cudaFree(ws->cell_own_cnt);
    ws->cell_start = ws->cell_fill = ws->cell_own = ws->scan_sums = ws->cell_own_cnt = nullptr;
    ws->cell_cap = 0;
    PANTEA_CUDA_TRY(cudaMalloc((void**)&ws->cell_start, 4 * (ncells + 1)));
    PANTEA_CUDA_TRY(cudaMalloc((void**)&ws->cell_fill, 4 * (ncells + 1)));
    PANTEA_CUDA_TRY(cudaMalloc((void**)&ws->cell_own, 4 * (ncells + 1)));
    PANTEA_CUDA_TRY(cudaMalloc((void**)&ws->scan_sums, 4 * 2 * (ncells / 2048 + 2)));
    PANTEA_CUDA_TRY(cudaMalloc((void**)&ws->cell_own_cnt, 4 * (ncells + 1)));
    PANTEA_CUDA_TRY(cudaMemset(ws->cell_own_cnt, 0, 4 * (ncells + 1)));
    ws->cell_cap = ncells;
    ws->scratch_clean = false;
    ws->skin_active = false;  // fresh (unzeroed) binning scratch: the next build is a forced one
    ++ws->arg_epoch;
    return PANTEA_OK;
}
}  // namespace pantea
