/* f16.h - C ABI of libf16_b200.so: the B200 (sm_100a) implementation of the `scores`
 * hot path of flake-it/flake16-framework.
 *
 * The reference has no FFI of its own: the path is Python calling scikit-learn /
 * imbalanced-learn estimators (experiment.py:73-100 construct them, :446-490 use them).
 * Each entry point below therefore names the reference call site (experiment.py:line) and the
 * third-party routine whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add.
 *
 * Conventions: every function returns 0 or a negative F16_ERR_* code and never throws;
 * f16_last_error() gives the thread-local message.  All *_dev pointers are device pointers
 * owned by the caller (any allocator; the Python layer uses torch tensors as containers).
 * Every call takes an explicit CUDA stream (cudaStream_t passed as void*) and is asynchronous
 * with respect to the host unless noted "synchronises".  Scratch memory is taken from the CUDA
 * stream-ordered pool (cudaMallocAsync) on that stream.  Opaque handles are freed only by their
 * *_free.  No global RNG state: seeds are arguments, mirroring the reference's
 * `random_state=0` re-seeding on every fit / fit_resample.
 */
#ifndef F16_H
#define F16_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define F16_OK 0
#define F16_ERR_INVALID (-1)
#define F16_ERR_CUDA (-2)
#define F16_ERR_OVERFLOW (-3)
#define F16_ERR_NOMEM (-4)

#define F16_KIND_DT 0 /* DecisionTreeClassifier(random_state=seed)   experiment.py:98 */
#define F16_KIND_RF 1 /* RandomForestClassifier(random_state=seed)   experiment.py:97 */
#define F16_KIND_ET 2 /* ExtraTreesClassifier(random_state=seed)     experiment.py:96 */

typedef struct f16_forest f16_forest;

const char* f16_last_error(void);
int f16_version(void);
/* Selects the device and keeps the stream-ordered memory pool from trimming. */
int f16_init(int device);
/* Number of kernels this library has launched (reset != 0 returns the count and zeroes it). */
long long f16_launch_count(int reset);
/* When on, f16_forest_fit brackets its tree-building kernel with CUDA events on the caller's
 * stream; f16_forest_build_ms() then waits for that kernel and returns its duration (ms). */
void f16_set_profiling(int on);
double f16_forest_build_ms(const f16_forest* forest);

/* ---- input: the parse step of load_feat_lab_proj (experiment.py:410-421), host only.
 * One-pass scanner for the tests.json wire format write_tests emits ({project: {test id: [req_runs,
 * label, f0..f15]}}); numbers are read with strtod, i.e. exactly as Python's json does.  Returns
 * F16_ERR_INVALID for anything outside that format (the Python layer then uses the json module).
 * f16_tests_copy: values_host float64 [rows][cols], proj_host int32 [rows] (index into the project
 * names, file order), names_host the names NUL separated (f16_tests_names_bytes bytes). */
typedef struct f16_tests f16_tests;
int f16_tests_parse(const char* path, f16_tests** out);
int64_t f16_tests_rows(const f16_tests* t);
int32_t f16_tests_cols(const f16_tests* t);
int32_t f16_tests_projects(const f16_tests* t);
int64_t f16_tests_names_bytes(const f16_tests* t);
int f16_tests_copy(const f16_tests* t, double* values_host, int32_t* proj_host, char* names_host);
void f16_tests_free(f16_tests* t);

/* ---- data staging ------------------------------------------------------------------
 * features[train] / features[test] (experiment.py:459-460) fused with the float32 cast that
 * sklearn applies in BaseForest.fit / predict.  X_dev: float64 [N][d] row-major; idx_dev: int64
 * row indices or NULL (identity); out: float32 [n_out][dp], dp = 8 if d <= 8 else 16, zero
 * padded - the row layout every tree kernel expects. */
int f16_gather_rows_f32(const double* X_dev, int32_t d, const int64_t* idx_dev, int64_t n_out,
                        float* out_dev, void* stream);
int f16_gather_rows_f64(const double* X_dev, int32_t d, const int64_t* idx_dev, int64_t n_out,
                        double* out_dev, void* stream);
int f16_gather_u8(const uint8_t* y_dev, const int64_t* idx_dev, int64_t n_out, uint8_t* out_dev, void* stream);
/* projects[test]  (experiment.py:460,477): project id of every test row. */
int f16_gather_i32(const int32_t* v_dev, const int64_t* idx_dev, int64_t n_out, int32_t* out_dev, void* stream);

/* ---- tree ensembles -------------------------------------------------------------------
 * Per-column argsort of a float32 row matrix ([n][dp]); sorted_idx_dev: int32 [d][n].
 * Replaces the per-node sort of sklearn's BestSplitter (tree/_partitioner.pyx:59-109). */
int f16_argsort_columns(const float* X_dev, int64_t n, int32_t d, int32_t* sorted_idx_dev, void* stream);

/* Host-only: per-tree seeds exactly as sklearn derives them from random_state=seed
 * (ensemble/_base.py:_set_random_states, tree/_splitter.pyx:155).  Arrays of n_trees. */
int f16_tree_seeds(uint32_t seed, int32_t kind, int32_t n_trees, uint32_t* tree_seed, uint32_t* rand_r_state);

/* RandomForest bootstrap: MT19937(tree_seed[t]).randint(0, n, n) -> bincount
 * (ensemble/_forest.py:94-112,150-156).  w_dev: uint8 [n_trees][(n+3)/4*4]. */
int f16_bootstrap_counts(const uint32_t* tree_seed_host, int32_t n_trees, int64_t n, uint8_t* w_dev, void* stream);

/* model.fit(features_train, labels_train)  (experiment.py:469).
 * X_dev float32 [n][dp] (from f16_gather_rows_f32), y_dev uint8 class index 0/1,
 * sorted_idx_dev from f16_argsort_columns (required for DT/RF, may be NULL for ET).
 * n_estimators is ignored for F16_KIND_DT.  max_features: sklearn's resolved value
 * (int(sqrt(d)) for the forests, d for the single tree). */
int f16_forest_fit(const float* X_dev, const uint8_t* y_dev, int64_t n, int32_t d,
                   const int32_t* sorted_idx_dev, int32_t kind, int32_t n_estimators,
                   int32_t max_features, uint32_t seed, void* stream, f16_forest** out);
/* The same fit with an explicit per-tree node capacity (0 = the worst case 2n - 1).  A tree that
 * outgrows the capacity makes f16_forest_status return F16_ERR_OVERFLOW; nothing is written out
 * of bounds.  The grid engine passes capacities measured on earlier fits (real trees of this path
 * hold 0.03 n - 0.15 n nodes) and refits with 0 on overflow. */
int f16_forest_fit_cap(const float* X_dev, const uint8_t* y_dev, int64_t n, int32_t d,
                       const int32_t* sorted_idx_dev, int32_t kind, int32_t n_estimators,
                       int32_t max_features, uint32_t seed, int64_t node_cap, void* stream, f16_forest** out);
/* model.predict(features_test)  (experiment.py:473): pred_dev uint8 [n] class index. */
int f16_forest_predict(const f16_forest* forest, const float* X_dev, int64_t n, uint8_t* pred_dev, void* stream);
/* Synchronises the stream; returns the device-side status of the fit (0 = ok). */
int f16_forest_status(const f16_forest* forest, void* stream);
int f16_forest_n_trees(const f16_forest* forest);
/* Largest node count among the forest's trees, as read by the last f16_forest_status (0 before). */
int f16_forest_max_nodes(const f16_forest* forest);
/* Synchronises. counts_host: int32 [n_trees]. */
int f16_forest_node_counts(const f16_forest* forest, int32_t* counts_host, void* stream);
/* Synchronises. One tree in sklearn's tree_ layout (children_left/right, feature, threshold,
 * impurity, n_node_samples, weighted_n_node_samples, value[n_nodes][2]) for parity tests. */
int f16_forest_export(const f16_forest* forest, int32_t tree, int64_t n_nodes, int64_t* left, int64_t* right,
                      int64_t* feature, double* threshold, double* impurity, int64_t* n_node_samples,
                      double* weighted_n_node_samples, double* value, void* stream);
void f16_forest_free(f16_forest* forest, void* stream);
/* shap.TreeExplainer(model).shap_values(features)[klass]  (get_shap, experiment.py:504-517): path-
 * dependent TreeSHAP of the fitted forest for every row of X_dev (float32 [n][dp] rows as for
 * predict), mean over the trees; phi_dev float64 [n][d].  Synchronises the stream once. */
int f16_forest_shap(const f16_forest* forest, const float* X_dev, int64_t n, int32_t klass, double* phi_dev,
                    void* stream);

/* ---- balancing: balancing.fit_resample(features_train, labels_train)  (experiment.py:463-466)
 * Exact float64 brute-force k-NN (NearestNeighbors(k).fit(A).kneighbors(Q)); idx_dev int32
 * [nq][k], ordered by (distance, index).  k <= 8, d <= 16.  col_order (HOST pointer, d ints, or
 * NULL) is the order in which coordinates are accumulated: highest-variance columns first make
 * the partial-distance early exit effective.  prefix_test selects the search strategy: 0 plain
 * float64 filter, 1 early exit on the two leading columns (raw features), 2 retired (= 0),
 * 3 tensor-core candidate filter (float16 x 3 split, mma.sync) followed by exact float64
 * selection - for centred data of moderate range (StandardScaler / PCA outputs); data that does
 * not fit float16 is detected on the device and searched exhaustively; 6 sweep over the rows sorted
 * by the leading column of col_order, outwards from the query until the gap in that column alone
 * exceeds the k-th best distance - for raw data whose variance sits in one column.  The result
 * depends on neither col_order nor prefix_test. */
int f16_knn(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, int32_t k,
            const int32_t* col_order, int32_t prefix_test, int32_t* idx_dev, void* stream);
/* Test hook for strategy 3.  Synchronises.  err_host receives the largest observed
 * |estimate - d^2| / (|q|^2 + |x|^2 + 1e-3) of the tensor-core distance estimate over all pairs
 * (the filter assumes 6e-5); n, nq <= 2^22, meant for small inputs. */
int f16_knn_tc_probe(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, float* err_host,
                     void* stream);
/* Same hook for the tcgen05 / TMEM implementation of the filter (strategy 3 uses it; strategy 5
 * forces the mma.sync implementation, strategy 4 forces the tcgen05 one whatever the size). */
int f16_knn_umma_probe(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, float* err_host,
                       void* stream);
/* y_new = full(n_new, minority) of SMOTE's output labels (a device memset). */
int f16_fill_u8(uint8_t* dst_dev, int32_t value, int64_t n, void* stream);
/* SMOTE._make_samples: X_new[j] = C[row] + steps[j] * (C[nn[row][1 + col]] - C[row]) with
 * row = sample_idx[j] / k, col = sample_idx[j] % k; nn_dev int32 [n_min][k + 1] from f16_knn. */
int f16_smote_generate(const double* C_dev, int64_t n_min, int32_t d, const int32_t* nn_dev, int32_t k,
                       const int64_t* sample_idx_dev, const double* steps_dev, int64_t n_new,
                       double* Xnew_dev, void* stream);
/* TomekLinks.is_tomek / EditedNearestNeighbours(kind_sel="all") decisions.  nn_dev int32 [n][kk]
 * (column 0 = the row itself), clean_mask bit c set <=> class c is cleaned; keep_dev uint8 [n]. */
int f16_tomek_keep(const int32_t* nn_dev, int32_t kk, const uint8_t* y_dev, int64_t n, int32_t clean_mask,
                   uint8_t* keep_dev, void* stream);
int f16_enn_keep(const int32_t* nn_dev, int32_t kk, const uint8_t* y_dev, int64_t n, int32_t clean_mask,
                 uint8_t* keep_dev, void* stream);
/* Row compaction: grouped = 0 keeps the original order (TomekLinks), 1 = class-0 rows then
 * class-1 rows (ENN).  src_index_dev (int64 [n], optional) receives the source row of each
 * output row; n_out_dev int64[2] = {rows kept, class-0 rows kept}. */
int f16_compact_rows(const double* X_dev, const uint8_t* y_dev, const uint8_t* keep_dev, int64_t n, int32_t d,
                     int32_t grouped, double* Xout_dev, uint8_t* yout_dev, int64_t* src_index_dev,
                     int64_t* n_out_dev, void* stream);

/* ---- figures: the 16 x 16 Spearman table of write_figures (experiment.py:661-663,
 * scipy.stats.spearmanr(features).correlation): average ranks (ties share the mean rank) per
 * column of X_dev (float64 [n][d] row-major), then the Pearson correlation of the ranks.
 * rho_dev float64 [d][d]. */
int f16_spearman(const double* X_dev, int64_t n, int32_t d, double* rho_dev, void* stream);

/* ---- scoring: the per-row loop of experiment.py:476-483.  counts_dev int64 [n_proj + 1][3]
 * (FP, FN, TP per project, last row = total), accumulated across calls (folds). */
int f16_confusion(const uint8_t* y_dev, const uint8_t* pred_dev, const int32_t* proj_dev, int64_t n,
                  int32_t n_proj, int64_t* counts_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
