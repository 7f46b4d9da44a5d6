/* ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use anything in oracle/).
 *
 * Plain-C, single-threaded restatement of the tree-ensemble arithmetic the reference's `scores`
 * path delegates to scikit-learn (reference call sites: experiment.py:96-98 construction,
 * :469 fit, :473 predict).  The dependency is **scikit-learn, pinned ==1.0.2** in the reference
 * (requirements.txt:31), not vendored under /root/reference; the algorithm below follows the
 * sources of the in-image scikit-learn 1.9.0 (SURVEY.md section 7.3 H5 discusses the skew):
 *
 *   forest seeds / bootstrap   sklearn/ensemble/_base.py:_set_random_states,
 *                              sklearn/ensemble/_forest.py:94-112 (_generate_sample_indices), :150-166
 *   numpy legacy RandomState   MT19937 init_genrand + masked-rejection randint
 *                              (numpy/random/src/distributions/distributions.c)
 *   builder                    sklearn/tree/_tree.pyx:139-336   (DepthFirstTreeBuilder.build)
 *   best splitter              sklearn/tree/_splitter.pyx:262-504, _partitioner.pyx:59-109,169-215
 *   random splitter            sklearn/tree/_splitter.pyx:507-736, _partitioner.pyx:129-167,217-246
 *   Gini                       sklearn/tree/_criterion.pyx:147-199, :622-687
 *   rand_r / rand_int / rand_uniform   sklearn/utils/_random.pxd:20-34, sklearn/tree/_utils.pyx:51-61
 *   predict                    sklearn/tree/_tree.pyx:954-996, sklearn/ensemble/_forest.py:704-717,882-967
 *
 * Pinned by tests/test_oracle_c_cpu.py against scikit-learn itself (node-for-node, bit-for-bit)
 * and against tests/golden/trees_n2000_seed16.npz (trees produced by the reference's own code path).
 * Build: `make -C oracle` -> oracle/_build/libtree_oracle.so   (gcc -O2 -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ MT19937 (numpy legacy) */
typedef struct { uint32_t mt[624]; int idx; } mt_t;
static void mt_seed(mt_t* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
static uint32_t mt_next(mt_t* s) {
    if (s->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}
static uint32_t mt_randint(mt_t* s, uint32_t hi) { /* RandomState.randint(0, hi) */
    uint32_t rng = hi - 1u, mask = rng, v;
    if (rng == 0) return 0;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    while ((v = (mt_next(s) & mask)) > rng) {}
    return v;
}

/* ------------------------------------------------------------------ sklearn our_rand_r */
static uint32_t rand_r32(uint32_t* s) {
    if (*s == 0) *s = 1;
    *s ^= (uint32_t)(*s << 13); *s ^= (uint32_t)(*s >> 17); *s ^= (uint32_t)(*s << 5);
    return *s % 0x80000000u;
}
static int rand_int(int lo, int hi, uint32_t* s) { return lo + (int)(rand_r32(s) % (uint32_t)(hi - lo)); }
static double rand_uniform(double lo, double hi, uint32_t* s) {
    return ((hi - lo) * (double)rand_r32(s) / 2147483647.0) + lo;
}

/* ------------------------------------------------------------------ tree storage (sklearn tree_ layout) */
typedef struct {
    int64_t cap, count;
    int64_t *left, *right, *feature, *n_node;
    double *threshold, *impurity, *w_node, *value; /* value[2*i + c] */
} tree_t;

static void tree_init(tree_t* t, int64_t cap) {
    t->cap = cap; t->count = 0;
    t->left = malloc(sizeof(int64_t) * cap); t->right = malloc(sizeof(int64_t) * cap);
    t->feature = malloc(sizeof(int64_t) * cap); t->n_node = malloc(sizeof(int64_t) * cap);
    t->threshold = malloc(sizeof(double) * cap); t->impurity = malloc(sizeof(double) * cap);
    t->w_node = malloc(sizeof(double) * cap); t->value = malloc(sizeof(double) * 2 * cap);
}
static void tree_free(tree_t* t) {
    free(t->left); free(t->right); free(t->feature); free(t->n_node);
    free(t->threshold); free(t->impurity); free(t->w_node); free(t->value);
}

static double gini(double a, double b, double w) {
    double sq = 0.0;
    sq += a * a; sq += b * b;
    return 1.0 - sq / (w * w);
}

typedef struct { int start, end, parent, is_left, n_const; double impurity; } rec_t;
typedef struct { float v; int s; } pair_t;
static int cmp_pair(const void* a, const void* b) {
    float x = ((const pair_t*)a)->v, y = ((const pair_t*)b)->v;
    return (x > y) - (x < y);
}

#define EPS 2.220446049250313e-16
#define FT 1e-7f

/* Grows one tree.  X: float32 [n][d] row-major.  w: sample weights (bootstrap counts) or NULL.
 * random_splitter: 0 best (RF, DT), 1 random (ET). */
static void grow_tree(const float* X, const uint8_t* y, int n, int d, const double* w, int max_features,
                      int random_splitter, uint32_t rand_r_state, tree_t* T) {
    int* samples = malloc(sizeof(int) * n);
    int ns = 0;
    double W_total = 0.0;
    for (int i = 0; i < n; i++) {
        if (!w || w[i] != 0.0) samples[ns++] = i;
        W_total += w ? w[i] : 1.0;
    }
    int features[64], const_feats[64];
    for (int f = 0; f < d; f++) { features[f] = f; const_feats[f] = 0; }
    pair_t* fv = malloc(sizeof(pair_t) * (ns > 0 ? ns : 1));
    rec_t* stack = malloc(sizeof(rec_t) * (2 * (size_t)ns + 4));
    int sp = 0, first = 1;
    uint32_t rs = rand_r_state;
    stack[sp++] = (rec_t){0, ns, -1, 0, 0, INFINITY};

    while (sp > 0) {
        rec_t r = stack[--sp];
        int start = r.start, end = r.end, nn = end - start;
        double t0 = 0.0, t1 = 0.0;
        for (int p = start; p < end; p++) {
            double ww = w ? w[samples[p]] : 1.0;
            if (y[samples[p]]) t1 += ww; else t0 += ww;
        }
        double wn = t0 + t1;
        double impurity = r.impurity;
        if (first) { impurity = gini(t0, t1, wn); first = 0; }
        int is_leaf = (nn < 2) || (impurity <= EPS);
        int best_f = -2, best_pos = end, n_total = r.n_const;
        double best_thr = -2.0, best_gl = 0, best_gr = 0, best_wl = 0, best_wr = 0;

        if (!is_leaf) {
            int f_i = d, n_visited = 0, n_found = 0, n_drawn = 0, n_known = r.n_const;
            double best_proxy = -INFINITY;
            while (f_i > n_total && (n_visited < max_features || n_visited <= n_found + n_drawn)) {
                n_visited++;
                int f_j = rand_int(n_drawn, f_i - n_found, &rs);
                if (f_j < n_known) {
                    int tmp = features[n_drawn]; features[n_drawn] = features[f_j]; features[f_j] = tmp;
                    n_drawn++;
                    continue;
                }
                f_j += n_found;
                int f = features[f_j];
                float vmin = INFINITY, vmax = -INFINITY;
                for (int p = start; p < end; p++) {
                    float v = X[(size_t)samples[p] * d + f];
                    fv[p - start].v = v; fv[p - start].s = samples[p];
                    if (v < vmin) vmin = v;
                    if (v > vmax) vmax = v;
                }
                if (vmax <= vmin + FT) {
                    features[f_j] = features[n_total]; features[n_total] = f;
                    n_found++; n_total++;
                    continue;
                }
                f_i--;
                { int tmp = features[f_i]; features[f_i] = features[f_j]; features[f_j] = tmp; }
                if (random_splitter) {
                    double thr = rand_uniform((double)vmin, (double)vmax, &rs);
                    if (thr == (double)vmax) thr = (double)vmin;
                    double l0 = 0.0, l1 = 0.0;
                    int nl = 0;
                    for (int p = 0; p < nn; p++) {
                        if ((double)fv[p].v <= thr) {
                            double ww = w ? w[fv[p].s] : 1.0;
                            if (y[fv[p].s]) l1 += ww; else l0 += ww;
                            nl++;
                        }
                    }
                    double wl = l0 + l1, wr = wn - wl;
                    double gl = gini(l0, l1, wl), gr = gini(t0 - l0, t1 - l1, wr);
                    double proxy = (-wr * gr) - wl * gl;
                    if (proxy > best_proxy) {
                        best_proxy = proxy; best_f = f; best_thr = thr; best_pos = start + nl;
                        best_gl = gl; best_gr = gr; best_wl = wl; best_wr = wr;
                    }
                } else {
                    qsort(fv, nn, sizeof(pair_t), cmp_pair);   /* any sort: ties are irrelevant */
                    double l0 = 0.0, l1 = 0.0;
                    int p = 0, done = 0;
                    while (1) {
                        int q = p + 1;
                        while (q < nn && fv[q].v <= fv[q - 1].v + FT) q++;
                        for (int k = p; k < q && k < nn; k++) {
                            double ww = w ? w[fv[k].s] : 1.0;
                            if (y[fv[k].s]) l1 += ww; else l0 += ww;
                        }
                        done = (q >= nn);
                        if (done) break;
                        p = q;
                        double wl = l0 + l1, wr = wn - wl;
                        double gl = gini(l0, l1, wl), gr = gini(t0 - l0, t1 - l1, wr);
                        double proxy = (-wr * gr) - wl * gl;
                        if (proxy > best_proxy) {
                            best_proxy = proxy; best_f = f; best_pos = start + p;
                            best_thr = (double)fv[p - 1].v / 2.0 + (double)fv[p].v / 2.0;
                            best_gl = gl; best_gr = gr; best_wl = wl; best_wr = wr;
                        }
                    }
                }
            }
            for (int i = 0; i < n_known; i++) features[i] = const_feats[i];
            for (int i = n_known; i < n_total; i++) const_feats[i] = features[i];
            if (best_pos >= end) is_leaf = 1;
            else {
                double improvement = (wn / W_total) * (impurity - (best_wr / wn * best_gr) - (best_wl / wn * best_gl));
                if (improvement + EPS < 0.0) is_leaf = 1;
            }
        }
        int64_t id = T->count++;
        T->left[id] = -1; T->right[id] = -1;
        T->feature[id] = is_leaf ? -2 : best_f;
        T->threshold[id] = is_leaf ? -2.0 : best_thr;
        T->impurity[id] = impurity;
        T->n_node[id] = nn; T->w_node[id] = wn;
        T->value[2 * id] = t0 / wn; T->value[2 * id + 1] = t1 / wn;
        if (r.parent >= 0) { if (r.is_left) T->left[r.parent] = id; else T->right[r.parent] = id; }
        if (!is_leaf) {
            /* partition samples[start:end) by X[s, f] <= thr (any order inside the halves) */
            int lo = start, hi = end;
            while (lo < hi) {
                if ((double)X[(size_t)samples[lo] * d + best_f] <= best_thr) lo++;
                else { hi--; int tmp = samples[lo]; samples[lo] = samples[hi]; samples[hi] = tmp; }
            }
            stack[sp++] = (rec_t){best_pos, end, (int)id, 0, n_total, best_gr};
            stack[sp++] = (rec_t){start, best_pos, (int)id, 1, n_total, best_gl};
        }
    }
    free(samples); free(fv); free(stack);
}

/* ------------------------------------------------------------------ public (ctypes) API */
typedef struct { int kind, n_trees, d; tree_t* trees; } forest_t;

/* kind: 0 DecisionTree, 1 RandomForest, 2 ExtraTrees (same codes as include/f16.h) */
forest_t* oracle_forest_fit(const float* X, const uint8_t* y, int n, int d, int kind, int n_estimators,
                            int max_features, uint32_t seed) {
    forest_t* F = malloc(sizeof(forest_t));
    F->kind = kind; F->d = d; F->n_trees = (kind == 0) ? 1 : n_estimators;
    F->trees = malloc(sizeof(tree_t) * F->n_trees);
    mt_t* rs = malloc(sizeof(mt_t));
    mt_t* ts = malloc(sizeof(mt_t));
    uint32_t* seeds = malloc(sizeof(uint32_t) * F->n_trees);
    if (kind == 0) seeds[0] = seed;
    else { mt_seed(rs, seed); for (int t = 0; t < F->n_trees; t++) seeds[t] = mt_randint(rs, 2147483647u); }
    double* w = malloc(sizeof(double) * n);
    for (int t = 0; t < F->n_trees; t++) {
        const double* wp = NULL;
        if (kind == 1) {                                   /* bootstrap counts as sample weights */
            memset(w, 0, sizeof(double) * n);
            mt_seed(ts, seeds[t]);
            for (int i = 0; i < n; i++) w[mt_randint(ts, (uint32_t)n)] += 1.0;
            wp = w;
        }
        mt_seed(ts, seeds[t]);
        uint32_t rr = mt_randint(ts, 2147483647u);
        tree_init(&F->trees[t], 2 * (int64_t)n + 1);
        grow_tree(X, y, n, d, wp, max_features, kind == 2, rr, &F->trees[t]);
    }
    free(w); free(seeds); free(rs); free(ts);
    return F;
}

int64_t oracle_tree_node_count(const forest_t* F, int t) { return F->trees[t].count; }

void oracle_tree_export(const forest_t* F, int t, int64_t* left, int64_t* right, int64_t* feature, double* threshold,
                        double* impurity, int64_t* n_node, double* w_node, double* value) {
    const tree_t* T = &F->trees[t];
    size_t c = (size_t)T->count;
    memcpy(left, T->left, 8 * c); memcpy(right, T->right, 8 * c); memcpy(feature, T->feature, 8 * c);
    memcpy(threshold, T->threshold, 8 * c); memcpy(impurity, T->impurity, 8 * c);
    memcpy(n_node, T->n_node, 8 * c); memcpy(w_node, T->w_node, 8 * c); memcpy(value, T->value, 16 * c);
}

/* predict: sum of leaf class fractions over trees in tree order, / n_trees, argmax (ties -> 0) */
void oracle_forest_predict(const forest_t* F, const float* X, int n, uint8_t* pred) {
    for (int i = 0; i < n; i++) {
        double p0 = 0.0, p1 = 0.0;
        for (int t = 0; t < F->n_trees; t++) {
            const tree_t* T = &F->trees[t];
            int64_t id = 0;
            while (T->left[id] != -1)
                id = ((double)X[(size_t)i * F->d + T->feature[id]] <= T->threshold[id]) ? T->left[id] : T->right[id];
            p0 += T->value[2 * id]; p1 += T->value[2 * id + 1];
        }
        if (F->n_trees > 1) { p0 /= (double)F->n_trees; p1 /= (double)F->n_trees; }
        pred[i] = p1 > p0;
    }
}

void oracle_forest_free(forest_t* F) {
    for (int t = 0; t < F->n_trees; t++) tree_free(&F->trees[t]);
    free(F->trees); free(F);
}

/* KAT helpers */
void oracle_randint(uint32_t seed, uint32_t hi, int n, uint32_t* out) {
    mt_t* s = malloc(sizeof(mt_t));
    mt_seed(s, seed);
    for (int i = 0; i < n; i++) out[i] = mt_randint(s, hi);
    free(s);
}
