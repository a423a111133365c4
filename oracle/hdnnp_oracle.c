/*
 * TEST INFRASTRUCTURE ONLY -- analytic CPU oracle (tier 2) for the HDNNP energy/force hot path.
 *
 * Plain C + OpenMP restatement of the reference algorithm with neighbour rows and analytic
 * central-role gradients, O(N * nbar^2).  It is validated against the faithful dense
 * restatement (oracle/dense_oracle.py, itself pinned by the reference's golden vectors) and is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * ONLY.  Nothing under pantea_b200/ links or calls it.
 *
 * Reference semantics followed (paths under /root/reference):
 *   minimum image, wrap            pantea/atoms/box.py:112-126
 *   distances, zero-vector guard   pantea/atoms/distance.py:63-78
 *   neighbour predicate            pantea/atoms/neighbor.py:102-107   (r <= rc) & (r > 0)
 *   cutoff functions               pantea/descriptors/acsf/cutoff.py:67-110  (fc uses r < rc)
 *   G1/G2                          pantea/descriptors/acsf/radial.py:39-61
 *   G3/G9 (r_shift ignored)        pantea/descriptors/acsf/angular.py:51-107
 *   radial / angular sums          pantea/descriptors/acsf/acsf.py:231-330  (x0.5 iff type_j == type_k,
 *                                  k == j excluded through r_jk > 0, r_jk = |pbc(d_ij - d_ik)|)
 *   scaler transforms              pantea/descriptors/scaler.py:206-246
 *   MLP + activations              pantea/models/nn/model.py:40-58, activation.py:7-60
 *   energy                         pantea/potentials/nnp/energy.py:23-66  (no atom_energy offset)
 *   force = -dE_i/dr_i, central    pantea/potentials/nnp/force.py:16-43
 *   velocity Verlet (no 1/m)       pantea/simulation/molecular_dynamics.py:16-77
 *   Berendsen, KE, T               pantea/simulation/thermostat.py:12-22, system.py:20-29
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_SF 256
#define ORC_MAX_WIDTH 256
#define ORC_MAX_LAYERS 16

typedef struct {
    int kind;        /* 1 G1, 2 G2, 3 G3, 9 G9 */
    int cutoff_type; /* 0 hard, 1 cos, 2 tanhu, 3 tanh, 4 exp, 5 poly1, 6 poly2 */
    int type_j;
    int type_k; /* 0 for radial */
    double r_cutoff, eta, r_shift, lambda0, zeta;
} orc_symfunc;

typedef struct {
    int central_type;
    int n_sf;
    const orc_symfunc *sf;
    /* scaler as x' = offset + slope * (x - shift); NULL pointers -> identity */
    const double *shift, *slope, *offset;
    /* MLP: sizes[n_layers + 1] (sizes[0] == n_sf), acts[n_layers], weights packed per layer:
       kernel [in,out] row-major then bias [out]; n_layers == 0 -> descriptor only */
    int n_layers;
    const int *sizes;
    const int *acts;
    const double *weights;
} orc_element;

static const double ORC_PI = 3.14159265358979323846;

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- geometry ------------------------------------------------------------------------- */
static inline double min_image(double dx, double box) { /* box.py:112-117 */
    if (dx > 0.5 * box) dx = dx - box;
    if (dx < -0.5 * box) dx = dx + box;
    return dx;
}

static inline double norm3(double x, double y, double z) { /* (x^2 + y^2) + z^2, uncontracted */
    if (x == 0.0 && y == 0.0 && z == 0.0) return 0.0;       /* distance.py:73-77 */
    return sqrt((x * x + y * y) + z * z);
}

void orc_wrap(double *pos, long n, const double *box) { /* box.py:123-126, floored remainder */
    for (long i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            double m = fmod(pos[3 * i + c], box[c]);
            if (m != 0.0 && ((m < 0.0) != (box[c] < 0.0))) m += box[c];
            pos[3 * i + c] = m;
        }
}

/* ---- cutoff function: value and derivative ------------------------------------------- */
static inline void cutoff_fn(int type, double r, double rc, double *fc, double *dfc) {
    if (!(r < rc)) { *fc = 0.0; *dfc = 0.0; return; } /* cutoff.py:67-72 */
    switch (type) {
    case 0: *fc = 1.0; *dfc = 0.0; break;
    case 1: { double a = ORC_PI * r / rc; *fc = 0.5 * (cos(a) + 1.0); *dfc = -0.5 * (ORC_PI / rc) * sin(a); break; }
    case 2:
    case 3: {
        double t = tanh(1.0 - r / rc), pre = 1.0;
        if (type == 3) { double e = exp(1.0), q = (e + 1.0 / e) / (e - 1.0 / e); pre = q * q * q; }
        *fc = pre * t * t * t;
        *dfc = pre * (-3.0 / rc) * t * t * (1.0 - t * t);
        break;
    }
    case 4: {
        double u = (r / rc) * (r / rc), om = 1.0 - u;
        double f = exp(1.0 - 1.0 / om);
        *fc = f; *dfc = -f * 2.0 * r / (rc * rc * om * om);
        break;
    }
    case 5: *fc = (2.0 * r - 3.0) * r * r + 1.0; *dfc = 6.0 * r * r - 6.0 * r; break;          /* raw r */
    case 6: *fc = ((15.0 - 6.0 * r) * r - 10.0) * r * r * r + 1.0;                               /* raw r */
            *dfc = -30.0 * r * r * r * r + 60.0 * r * r * r - 30.0 * r * r; break;
    default: *fc = 0.0; *dfc = 0.0;
    }
}

/* ---- activation: value and derivative (activation.py:7-60) --------------------------- */
static inline void act_fn(int a, double x, double *y, double *dy) {
    switch (a) {
    case 0: *y = x; *dy = 1.0; break;
    case 1: { double t = tanh(x); *y = t; *dy = 1.0 - t * t; break; }
    case 2: { double s = 1.0 / (1.0 + exp(-x)); *y = s; *dy = s * (1.0 - s); break; }
    case 3: { double s = 1.0 / (1.0 + exp(-x)); *y = (x > 0 ? x : 0.0) + log1p(exp(-fabs(x))); *dy = s; break; }
    case 4: *y = x > 0 ? x : 0.0; *dy = x > 0 ? 1.0 : 0.0; break;
    case 5: { double g = exp(-0.5 * x * x); *y = g; *dy = -x * g; break; }
    case 6: *y = cos(x); *dy = -sin(x); break;
    case 7: { double e = exp(-x); *y = e; *dy = -e; break; }
    case 8: *y = x * x; *dy = 2.0 * x; break;
    default: *y = x; *dy = 1.0;
    }
}

/* ---- neighbour rows ------------------------------------------------------------------- */
typedef struct { double dx, dy, dz, r; int idx, type; } orc_nbr;

/* all neighbours of a centre at p (structure atoms [0,n)), ascending index; returns count */
static int gather_neighbors(const double *p, const double *pos, const int *types, long n, const double *box,
                            double rc, orc_nbr *out, int cap) {
    int cnt = 0;
    for (long j = 0; j < n; ++j) {
        double dx = p[0] - pos[3 * j], dy = p[1] - pos[3 * j + 1], dz = p[2] - pos[3 * j + 2];
        if (box) { dx = min_image(dx, box[0]); dy = min_image(dy, box[1]); dz = min_image(dz, box[2]); }
        double r = norm3(dx, dy, dz);
        if (r <= rc && r > 0.0) { /* neighbor.py:107 */
            if (cnt < cap) { out[cnt].dx = dx; out[cnt].dy = dy; out[cnt].dz = dz; out[cnt].r = r;
                             out[cnt].idx = (int)j; out[cnt].type = types[j]; }
            ++cnt;
        }
    }
    return cnt;
}


/* ---- optional cell list (same neighbour sets and the same ascending order as the all-pairs scan) ------ */
typedef struct { int n[3]; double inv[3]; int *head, *next; int active; } orc_cells;

static int orc_cell_coord(double x, double inv, int n) {
    int c = (int)floor(x * inv);
    c %= n; if (c < 0) c += n;
    return c;
}

static int g_use_cells = 1;
void orc_set_use_cells(int flag) { g_use_cells = flag; }

static void orc_cells_build(orc_cells *cl, const double *pos, long n, const double *box, double rc) {
    cl->active = 0; cl->head = cl->next = NULL;
    if (!box || n < 64 || !g_use_cells) return;
    double w = rc * (1.0 + 1e-7);
    for (int k = 0; k < 3; ++k) {
        cl->n[k] = (int)floor(box[k] / w);
        if (cl->n[k] < 3) return;
        if (cl->n[k] > 256) cl->n[k] = 256;
        cl->inv[k] = cl->n[k] / box[k];
    }
    long nc = (long)cl->n[0] * cl->n[1] * cl->n[2];
    cl->head = (int *)malloc(sizeof(int) * (size_t)nc);
    cl->next = (int *)malloc(sizeof(int) * (size_t)n);
    for (long c = 0; c < nc; ++c) cl->head[c] = -1;
    for (long i = n - 1; i >= 0; --i) { /* reverse insertion -> ascending index inside a cell */
        int cx = orc_cell_coord(pos[3 * i], cl->inv[0], cl->n[0]), cy = orc_cell_coord(pos[3 * i + 1], cl->inv[1], cl->n[1]),
            cz = orc_cell_coord(pos[3 * i + 2], cl->inv[2], cl->n[2]);
        long c = ((long)cz * cl->n[1] + cy) * cl->n[0] + cx;
        cl->next[i] = cl->head[c]; cl->head[c] = (int)i;
    }
    cl->active = 1;
}

static void orc_cells_free(orc_cells *cl) { free(cl->head); free(cl->next); cl->head = cl->next = NULL; cl->active = 0; }

static int nbr_cmp(const void *a, const void *b) { return ((const orc_nbr *)a)->idx - ((const orc_nbr *)b)->idx; }

static int gather_neighbors_cells(const orc_cells *cl, const double *p, const double *pos, const int *types,
                                  const double *box, double rc, orc_nbr *out, int cap) {
    int cnt = 0;
    int c0[3] = {orc_cell_coord(p[0], cl->inv[0], cl->n[0]), orc_cell_coord(p[1], cl->inv[1], cl->n[1]),
                 orc_cell_coord(p[2], cl->inv[2], cl->n[2])};
    for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx_ = -1; dx_ <= 1; ++dx_) {
        int x = (c0[0] + dx_ + cl->n[0]) % cl->n[0], y = (c0[1] + dy + cl->n[1]) % cl->n[1], z = (c0[2] + dz + cl->n[2]) % cl->n[2];
        for (int j = cl->head[((long)z * cl->n[1] + y) * cl->n[0] + x]; j >= 0; j = cl->next[j]) {
            double dx = p[0] - pos[3 * j], dy2 = p[1] - pos[3 * j + 1], dz2 = p[2] - pos[3 * j + 2];
            dx = min_image(dx, box[0]); dy2 = min_image(dy2, box[1]); dz2 = min_image(dz2, box[2]);
            double r = norm3(dx, dy2, dz2);
            if (r <= rc && r > 0.0) {
                if (cnt < cap) { out[cnt].dx = dx; out[cnt].dy = dy2; out[cnt].dz = dz2; out[cnt].r = r;
                                 out[cnt].idx = j; out[cnt].type = types[j]; }
                ++cnt;
            }
        }
    }
    qsort(out, (size_t)(cnt < cap ? cnt : cap), sizeof(orc_nbr), nbr_cmp);
    return cnt;
}

/* Neighbour lists as CSR with ascending columns.  Pass col == NULL to only count.
   Returns total number of neighbours. */
long orc_neighbors(const double *pos, const int *types, long n, const double *box, double rc, long *row_ptr,
                   int *col, long cap) {
    int *cnt = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        int c = 0;
        for (long j = 0; j < n; ++j) {
            double dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
            if (box) { dx = min_image(dx, box[0]); dy = min_image(dy, box[1]); dz = min_image(dz, box[2]); }
            double r = norm3(dx, dy, dz);
            if (r <= rc && r > 0.0) ++c;
        }
        cnt[i] = c;
    }
    row_ptr[0] = 0;
    for (long i = 0; i < n; ++i) row_ptr[i + 1] = row_ptr[i] + cnt[i];
    long total = row_ptr[n];
    free(cnt);
    if (!col || total > cap) return total;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        long o = row_ptr[i];
        for (long j = 0; j < n; ++j) {
            double dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
            if (box) { dx = min_image(dx, box[0]); dy = min_image(dy, box[1]); dz = min_image(dz, box[2]); }
            double r = norm3(dx, dy, dz);
            if (r <= rc && r > 0.0) col[o++] = (int)j;
        }
    }
    return total;
}

/* dense distance matrix rows for `calculate_distances` parity (distance.py:63-105) */
void orc_distances(const double *pos, long n, const double *box, double *out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i)
        for (long j = 0; j < n; ++j) {
            double dx = pos[3 * i] - pos[3 * j], dy = pos[3 * i + 1] - pos[3 * j + 1], dz = pos[3 * i + 2] - pos[3 * j + 2];
            if (box) { dx = min_image(dx, box[0]); dy = min_image(dy, box[1]); dz = min_image(dz, box[2]); }
            out[i * n + j] = norm3(dx, dy, dz);
        }
}

/* ---- descriptor of one centre: values G[n_sf] and central gradient dG[n_sf][3] ---------- */
static inline double ipow_or_pow(double base, double e) {
    double ie = floor(e);
    if (ie == e && e >= 0.0 && e <= 64.0) {
        double r = 1.0; int n = (int)ie;
        for (int i = 0; i < n; ++i) r *= base;
        return r;
    }
    return pow(base, e);
}

static void acsf_one(const orc_element *el, const orc_nbr *nb, int nn, const double *box, double *G, double *dG,
                     double *scratch /* >= 2*nn doubles */) {
    double *fcn = scratch, *dfcn = scratch + nn;
    for (int s = 0; s < el->n_sf; ++s) {
        const orc_symfunc *sf = &el->sf[s];
        double g = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
        if (sf->kind == 1 || sf->kind == 2) {
            for (int a = 0; a < nn; ++a) {
                if (nb[a].type != sf->type_j) continue;
                double r = nb[a].r, fc, dfc;
                cutoff_fn(sf->cutoff_type, r, sf->r_cutoff, &fc, &dfc);
                double val, dval;
                if (sf->kind == 1) { val = fc; dval = dfc; }
                else {
                    double dr = r - sf->r_shift, ex = exp(-sf->eta * dr * dr);
                    val = ex * fc; dval = ex * (dfc - 2.0 * sf->eta * dr * fc);
                }
                g += val;
                double s_r = dval / r;
                gx += s_r * nb[a].dx; gy += s_r * nb[a].dy; gz += s_r * nb[a].dz;
            }
        } else {
            const int same = sf->type_j == sf->type_k;
            const double pref = pow(2.0, 1.0 - sf->zeta), lam = sf->lambda0, zeta = sf->zeta, eta = sf->eta;
            for (int a = 0; a < nn; ++a) cutoff_fn(sf->cutoff_type, nb[a].r, sf->r_cutoff, &fcn[a], &dfcn[a]);
            for (int a = 0; a < nn; ++a) {
                if (nb[a].type != sf->type_j) continue;
                double rj = nb[a].r, fcj = fcn[a], dfcj = dfcn[a];
                if (fcj == 0.0 && dfcj == 0.0) continue;
                for (int b = (same ? a + 1 : 0); b < nn; ++b) {
                    if (nb[b].type != sf->type_k) continue;
                    double rk = nb[b].r, fck = fcn[b], dfck = dfcn[b];
                    if (fck == 0.0 && dfck == 0.0) continue;
                    double ex = nb[a].dx - nb[b].dx, ey = nb[a].dy - nb[b].dy, ez = nb[a].dz - nb[b].dz;
                    if (box) { ex = min_image(ex, box[0]); ey = min_image(ey, box[1]); ez = min_image(ez, box[2]); }
                    double rjk = norm3(ex, ey, ez);
                    if (!(rjk > 0.0)) continue; /* acsf.py:325 */
                    double fcjk = 1.0, r2 = rj * rj + rk * rk;
                    if (sf->kind == 3) {
                        double dummy;
                        cutoff_fn(sf->cutoff_type, rjk, sf->r_cutoff, &fcjk, &dummy);
                        if (fcjk == 0.0) continue;
                        r2 += rjk * rjk;
                    }
                    double inv = 1.0 / (rj * rk);
                    double cost = (nb[a].dx * nb[b].dx + nb[a].dy * nb[b].dy + nb[a].dz * nb[b].dz) * inv;
                    double base = 1.0 + lam * cost;
                    double pw1 = ipow_or_pow(base, zeta - 1.0); /* base^(zeta-1) */
                    double ang = pref * pw1 * base, e = exp(-eta * r2);
                    double fprod = fcj * fck * fcjk;
                    double T = ang * e * fprod;
                    double Tc = pref * zeta * lam * pw1 * e * fprod;
                    double Tij = ang * e * (-2.0 * eta * rj * fprod + dfcj * fck * fcjk);
                    double Tik = ang * e * (-2.0 * eta * rk * fprod + fcj * dfck * fcjk);
                    double Aj = Tc * (inv - cost / (rj * rj)) + Tij / rj;
                    double Ak = Tc * (inv - cost / (rk * rk)) + Tik / rk;
                    g += T;
                    gx += Aj * nb[a].dx + Ak * nb[b].dx;
                    gy += Aj * nb[a].dy + Ak * nb[b].dy;
                    gz += Aj * nb[a].dz + Ak * nb[b].dz;
                }
            }
            /* unordered pairs already equal 0.5 * (ordered sum) when type_j == type_k (acsf.py:285-290) */
        }
        G[s] = g;
        if (dG) { dG[3 * s] = gx; dG[3 * s + 1] = gy; dG[3 * s + 2] = gz; }
    }
}

static double max_cutoff(const orc_element *el) {
    double rc = 0.0;
    for (int s = 0; s < el->n_sf; ++s) if (el->sf[s].r_cutoff > rc) rc = el->sf[s].r_cutoff;
    return rc;
}

/* Descriptor values/gradients of `el` for the centres `centres[n_c]` (any atom type, like
   ACSF.grad with atom_index=None).  G [n_c, n_sf], dG [n_c, n_sf, 3] (may be NULL). */
int orc_acsf(const orc_element *el, const double *pos, const int *types, long n, const double *box,
             const int *centres, long n_c, double *G, double *dG) {
    if (el->n_sf > ORC_MAX_SF) return -1;
    const double rc = max_cutoff(el);
    int err = 0;
#pragma omp parallel
    {
        orc_nbr *nb = (orc_nbr *)malloc(sizeof(orc_nbr) * (size_t)(n > 0 ? n : 1));
        double *scratch = (double *)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 4)
        for (long c = 0; c < n_c; ++c) {
            long i = centres ? centres[c] : c;
            int nn = gather_neighbors(pos + 3 * i, pos, types, n, box, rc, nb, (int)n);
            acsf_one(el, nb, nn, box, G + c * el->n_sf, dG ? dG + c * el->n_sf * 3 : NULL, scratch);
        }
        free(nb); free(scratch);
    }
    return err;
}

/* ---- scaler + MLP: energy and dE/dG -------------------------------------------------- */
static int mlp_one(const orc_element *el, const double *G, double *E, double *dEdG) {
    double h[ORC_MAX_LAYERS + 1][ORC_MAX_WIDTH], dact[ORC_MAX_LAYERS][ORC_MAX_WIDTH];
    const int L = el->n_layers;
    if (L > ORC_MAX_LAYERS) return -1;
    for (int s = 0; s < el->n_sf; ++s)
        h[0][s] = el->slope ? el->offset[s] + el->slope[s] * (G[s] - el->shift[s]) : G[s];
    const double *w = el->weights;
    const double *wl[ORC_MAX_LAYERS];
    for (int l = 0; l < L; ++l) {
        int ni = el->sizes[l], no = el->sizes[l + 1];
        if (ni > ORC_MAX_WIDTH || no > ORC_MAX_WIDTH) return -1;
        wl[l] = w;
        const double *b = w + (size_t)ni * no;
        for (int o = 0; o < no; ++o) {
            double z = 0.0;
            for (int i = 0; i < ni; ++i) z += h[l][i] * w[(size_t)i * no + o];
            z += b[o];
            act_fn(el->acts[l], z, &h[l + 1][o], &dact[l][o]);
        }
        w = b + no;
    }
    *E = h[L][0];
    /* back-substitution for dE/dh_0 */
    double gcur[ORC_MAX_WIDTH], gprev[ORC_MAX_WIDTH];
    for (int o = 0; o < el->sizes[L]; ++o) gcur[o] = (o == 0) ? 1.0 : 0.0;
    for (int l = L - 1; l >= 0; --l) {
        int ni = el->sizes[l], no = el->sizes[l + 1];
        for (int i = 0; i < ni; ++i) {
            double a = 0.0;
            for (int o = 0; o < no; ++o) a += wl[l][(size_t)i * no + o] * (gcur[o] * dact[l][o]);
            gprev[i] = a;
        }
        memcpy(gcur, gprev, sizeof(double) * (size_t)ni);
    }
    for (int s = 0; s < el->n_sf; ++s) dEdG[s] = gcur[s] * (el->slope ? el->slope[s] : 1.0);
    return 0;
}

static const orc_element *find_element(const orc_element *els, int n_el, int type) {
    for (int e = 0; e < n_el; ++e) if (els[e].central_type == type) return &els[e];
    return NULL;
}

/* Energies and central-role forces for atoms [begin,end) of a structure of n atoms.
   e_atom [n] and forces [n,3] are written only for that range (others untouched);
   returns the partial energy sum through *e_sum. */
int orc_energy_forces_range(const orc_element *els, int n_el, const double *pos, const int *types, long n,
                            const double *box, long begin, long end, double *e_atom, double *forces, double *e_sum) {
    double rc = 0.0;
    for (int e = 0; e < n_el; ++e) { double r = max_cutoff(&els[e]); if (r > rc) rc = r; }
    int err = 0;
    double total = 0.0;
    orc_cells cells;
    orc_cells_build(&cells, pos, n, box, rc);
#pragma omp parallel
    {
        orc_nbr *nb = (orc_nbr *)malloc(sizeof(orc_nbr) * (size_t)(n > 0 ? n : 1));
        double *scratch = (double *)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
        double G[ORC_MAX_SF], dG[ORC_MAX_SF * 3], w[ORC_MAX_SF];
#pragma omp for schedule(dynamic, 4) reduction(+ : total)
        for (long i = begin; i < end; ++i) {
            const orc_element *el = find_element(els, n_el, types[i]);
            double E = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
            if (el && el->n_sf <= ORC_MAX_SF && el->n_layers > 0) {
                int nn = cells.active ? gather_neighbors_cells(&cells, pos + 3 * i, pos, types, box, rc, nb, (int)n)
                                      : gather_neighbors(pos + 3 * i, pos, types, n, box, rc, nb, (int)n);
                acsf_one(el, nb, nn, box, G, forces ? dG : NULL, scratch);
                if (mlp_one(el, G, &E, w) != 0) { err = -1; }
                if (forces)
                    for (int s = 0; s < el->n_sf; ++s) {
                        fx -= w[s] * dG[3 * s]; fy -= w[s] * dG[3 * s + 1]; fz -= w[s] * dG[3 * s + 2];
                    }
            }
            if (e_atom) e_atom[i] = E;
            if (forces) { forces[3 * i] = fx; forces[3 * i + 1] = fy; forces[3 * i + 2] = fz; }
            total += E;
        }
        free(nb); free(scratch);
    }
    orc_cells_free(&cells);
    if (e_sum) *e_sum = total;
    return err;
}

int orc_energy_forces(const orc_element *els, int n_el, const double *pos, const int *types, long n,
                      const double *box, double *e_atom, double *forces, double *e_total) {
    int rc = orc_energy_forces_range(els, n_el, pos, types, n, box, 0, n, e_atom, forces, NULL);
    /* deterministic total: plain ascending sum of the per-atom energies */
    if (e_total) {
        if (e_atom) { double t = 0.0; for (long i = 0; i < n; ++i) t += e_atom[i]; *e_total = t; }
        else { double t; orc_energy_forces_range(els, n_el, pos, types, n, box, 0, n, NULL, NULL, &t); *e_total = t; }
    }
    return rc;
}

/* ---- full force (extension, NOT a reference mode): F = -dE/dr through both roles ------------------------
   The oracle of PANTEA_FORCE_FULL at sizes the dense autograd restatement cannot reach; itself checked against
   that restatement in tests/test_oracle_golden.py.  With w_s = dE_i/dG_is, centre i contributes
   g_in = sum_s w_s dG_is/d(d_in) to itself (F_i -= g_in) and to neighbour n (F_n += g_in); for a G3 triplet
      dT/dd_ij = T_c (u_k - c u_j)/r_j + T_j u_j + T_jk u_jk,
                            dT/dd_ik = T_c (u_j - c u_k)/r_k + T_k u_k - T_jk u_jk.
   fi[3] accumulates -sum_n g_in, fn[nn][3] the g_in per neighbour. */
static void acsf_scatter_one(const orc_element *el, const orc_nbr *nb, int nn, const double *box, const double *w,
                             double *fi, double *fn, double *scratch /* >= 2*nn doubles */) {
    double *fcn = scratch, *dfcn = scratch + nn;
    for (int a = 0; a < 3 * nn; ++a) fn[a] = 0.0;
    for (int s = 0; s < el->n_sf; ++s) {
        const orc_symfunc *sf = &el->sf[s];
        if (sf->kind == 1 || sf->kind == 2) {
            for (int a = 0; a < nn; ++a) {
                if (nb[a].type != sf->type_j) continue;
                double r = nb[a].r, fc, dfc, dval;
                cutoff_fn(sf->cutoff_type, r, sf->r_cutoff, &fc, &dfc);
                if (sf->kind == 1) dval = dfc;
                else {
                    double dr = r - sf->r_shift, ex = exp(-sf->eta * dr * dr);
                    dval = ex * (dfc - 2.0 * sf->eta * dr * fc);
                }
                double c = w[s] * dval / r;
                fn[3 * a] += c * nb[a].dx; fn[3 * a + 1] += c * nb[a].dy; fn[3 * a + 2] += c * nb[a].dz;
            }
            continue;
        }
        const int same = sf->type_j == sf->type_k;
        const double pref = pow(2.0, 1.0 - sf->zeta), lam = sf->lambda0, zeta = sf->zeta, eta = sf->eta;
        for (int a = 0; a < nn; ++a) cutoff_fn(sf->cutoff_type, nb[a].r, sf->r_cutoff, &fcn[a], &dfcn[a]);
        for (int a = 0; a < nn; ++a) {
            if (nb[a].type != sf->type_j) continue;
            double rj = nb[a].r, fcj = fcn[a], dfcj = dfcn[a];
            if (fcj == 0.0 && dfcj == 0.0) continue;
            for (int b = (same ? a + 1 : 0); b < nn; ++b) {
                if (nb[b].type != sf->type_k) continue;
                double rk = nb[b].r, fck = fcn[b], dfck = dfcn[b];
                if (fck == 0.0 && dfck == 0.0) continue;
                double ex = nb[a].dx - nb[b].dx, ey = nb[a].dy - nb[b].dy, ez = nb[a].dz - nb[b].dz;
                if (box) { ex = min_image(ex, box[0]); ey = min_image(ey, box[1]); ez = min_image(ez, box[2]); }
                double rjk = norm3(ex, ey, ez);
                if (!(rjk > 0.0)) continue; /* acsf.py:325 */
                double fcjk = 1.0, dfcjk = 0.0, r2 = rj * rj + rk * rk;
                if (sf->kind == 3) {
                    cutoff_fn(sf->cutoff_type, rjk, sf->r_cutoff, &fcjk, &dfcjk);
                    if (fcjk == 0.0 && dfcjk == 0.0) continue;
                    r2 += rjk * rjk;
                }
                double cost = (nb[a].dx * nb[b].dx + nb[a].dy * nb[b].dy + nb[a].dz * nb[b].dz) / (rj * rk);
                double base = 1.0 + lam * cost;
                double pw1 = ipow_or_pow(base, zeta - 1.0);
                double ang = pref * pw1 * base, e = exp(-eta * r2);
                double Tc = w[s] * pref * zeta * lam * pw1 * e * fcj * fck * fcjk;
                double Tj = w[s] * ang * e * fck * fcjk * (dfcj - 2.0 * eta * rj * fcj);
                double Tk = w[s] * ang * e * fcj * fcjk * (dfck - 2.0 * eta * rk * fck);
                double Tjk = sf->kind == 3 ? w[s] * ang * e * fcj * fck * (dfcjk - 2.0 * eta * rjk * fcjk) : 0.0;
                double a1 = (Tj - Tc * cost / rj) / rj, a2 = Tc / (rj * rk), a3 = Tjk / rjk;
                double b1 = (Tk - Tc * cost / rk) / rk;
                double gjx = a1 * nb[a].dx + a2 * nb[b].dx + a3 * ex, gjy = a1 * nb[a].dy + a2 * nb[b].dy + a3 * ey,
                       gjz = a1 * nb[a].dz + a2 * nb[b].dz + a3 * ez;
                double gkx = b1 * nb[b].dx + a2 * nb[a].dx - a3 * ex, gky = b1 * nb[b].dy + a2 * nb[a].dy - a3 * ey,
                       gkz = b1 * nb[b].dz + a2 * nb[a].dz - a3 * ez;
                fn[3 * a] += gjx; fn[3 * a + 1] += gjy; fn[3 * a + 2] += gjz;
                fn[3 * b] += gkx; fn[3 * b + 1] += gky; fn[3 * b + 2] += gkz;
            }
        }
    }
    fi[0] = fi[1] = fi[2] = 0.0;
    for (int a = 0; a < nn; ++a) { fi[0] -= fn[3 * a]; fi[1] -= fn[3 * a + 1]; fi[2] -= fn[3 * a + 2]; }
}

/* e_atom [n], forces [n,3] (full force), e_total; thread-private force arrays summed in thread order */
int orc_energy_full_forces(const orc_element *els, int n_el, const double *pos, const int *types, long n,
                           const double *box, double *e_atom, double *forces, double *e_total) {
    double rc = 0.0;
    for (int e = 0; e < n_el; ++e) { double r = max_cutoff(&els[e]); if (r > rc) rc = r; }
    int err = 0;
    orc_cells cells;
    orc_cells_build(&cells, pos, n, box, rc);
    int n_threads = 1;
#ifdef _OPENMP
    n_threads = omp_get_max_threads();
#endif
    double *priv = (double *)calloc((size_t)n_threads * 3 * (size_t)(n > 0 ? n : 1), sizeof(double));
#pragma omp parallel num_threads(n_threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *mine = priv + (size_t)tid * 3 * (size_t)n;
        orc_nbr *nb = (orc_nbr *)malloc(sizeof(orc_nbr) * (size_t)(n > 0 ? n : 1));
        double *scratch = (double *)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
        double *fn = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
        double G[ORC_MAX_SF], w[ORC_MAX_SF];
#pragma omp for schedule(static)
        for (long i = 0; i < n; ++i) {
            const orc_element *el = find_element(els, n_el, types[i]);
            double E = 0.0;
            if (el && el->n_sf <= ORC_MAX_SF && el->n_layers > 0) {
                int nn = cells.active ? gather_neighbors_cells(&cells, pos + 3 * i, pos, types, box, rc, nb, (int)n)
                                      : gather_neighbors(pos + 3 * i, pos, types, n, box, rc, nb, (int)n);
                acsf_one(el, nb, nn, box, G, NULL, scratch);
                if (mlp_one(el, G, &E, w) != 0) err = -1;
                double fi[3];
                acsf_scatter_one(el, nb, nn, box, w, fi, fn, scratch);
                mine[3 * i] += fi[0]; mine[3 * i + 1] += fi[1]; mine[3 * i + 2] += fi[2];
                for (int a = 0; a < nn; ++a) {
                    const long j = nb[a].idx;
                    mine[3 * j] += fn[3 * a]; mine[3 * j + 1] += fn[3 * a + 1]; mine[3 * j + 2] += fn[3 * a + 2];
                }
            }
            if (e_atom) e_atom[i] = E;
        }
        free(nb); free(scratch); free(fn);
    }
    if (forces)
        for (long k = 0; k < 3 * n; ++k) {
            double t = 0.0;
            for (int th = 0; th < n_threads; ++th) t += priv[(size_t)th * 3 * (size_t)n + k];
            forces[k] = t;
        }
    free(priv);
    orc_cells_free(&cells);
    if (e_total && e_atom) { double t = 0.0; for (long i = 0; i < n; ++i) t += e_atom[i]; *e_total = t; }
    return err;
}

/* ---- MD driver: velocity Verlet without mass, optional Berendsen --------------------- */
/* scalars_out [n_steps + 1, 3] = (E_pot, E_kin, T) recorded at step 0..n_steps when not NULL.
   tau <= 0 disables the thermostat.  pos/vel/forces are updated in place; forces must hold
   F(pos) on entry (like System.__post_init__, system.py:79-82). */
/* extension switch (not a reference mode): accelerations F/m instead of F in both half-steps -- the oracle of
   pantea_md_params.mass_scaled */
static int g_mass_scaled = 0;
void orc_set_mass_scaled(int flag) { g_mass_scaled = flag; }

int orc_md_run(const orc_element *els, int n_el, double *pos, double *vel, double *forces, const double *mass,
               const int *types, long n, const double *box, double dt, long n_steps, double t_target, double tau,
               double kb, double *scalars_out) {
    double *fnew = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *eat = (double *)malloc(sizeof(double) * (size_t)n);
    int rc = 0;
    for (long step = 0; step <= n_steps; ++step) {
        if (scalars_out) {
            double ep = 0.0, ke = 0.0;
            if (step == 0) rc |= orc_energy_forces(els, n_el, pos, types, n, box, eat, NULL, &ep);
            else { for (long i = 0; i < n; ++i) ep += eat[i]; }
            for (long i = 0; i < n; ++i)
                ke += mass[i] * vel[3 * i] * vel[3 * i] + mass[i] * vel[3 * i + 1] * vel[3 * i + 1] +
                      mass[i] * vel[3 * i + 2] * vel[3 * i + 2];
            ke *= 0.5;
            scalars_out[3 * step] = ep; scalars_out[3 * step + 1] = ke;
            scalars_out[3 * step + 2] = 2.0 * ke / (3.0 * (double)n * kb);
        }
        if (step == n_steps) break;
        if (g_mass_scaled)
            for (long i = 0; i < 3 * n; ++i) pos[i] = pos[i] + vel[i] * dt + 0.5 * (forces[i] / mass[i / 3]) * dt * dt;
        else
            for (long i = 0; i < 3 * n; ++i) pos[i] = pos[i] + vel[i] * dt + 0.5 * forces[i] * dt * dt;
        if (box) orc_wrap(pos, n, box);
        rc |= orc_energy_forces(els, n_el, pos, types, n, box, eat, fnew, NULL);
        if (g_mass_scaled)
            for (long i = 0; i < 3 * n; ++i) vel[i] = vel[i] + 0.5 * (forces[i] / mass[i / 3] + fnew[i] / mass[i / 3]) * dt;
        else
            for (long i = 0; i < 3 * n; ++i) vel[i] = vel[i] + 0.5 * (forces[i] + fnew[i]) * dt;
        memcpy(forces, fnew, sizeof(double) * 3 * (size_t)n);
        if (tau > 0.0) {
            double ke = 0.0;
            for (long i = 0; i < n; ++i)
                ke += mass[i] * vel[3 * i] * vel[3 * i] + mass[i] * vel[3 * i + 1] * vel[3 * i + 1] +
                      mass[i] * vel[3 * i + 2] * vel[3 * i + 2];
            double T = 2.0 * (0.5 * ke) / (3.0 * (double)n * kb);
            double s = 1.0 / sqrt(1.0 + (dt / tau) * (T / t_target - 1.0));
            for (long i = 0; i < 3 * n; ++i) vel[i] *= s;
        }
    }
    free(fnew); free(eat);
    return rc;
}
