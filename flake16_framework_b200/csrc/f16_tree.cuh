// Shared declarations for the tree builder / predictor kernels.
#pragma once
#include "f16_common.cuh"

// One tree node, 32 bytes (two 16-byte loads when walking).  Nodes are numbered in the
// order sklearn's DepthFirstTreeBuilder pops them (sklearn/tree/_tree.pyx:139-336):
// pre-order, left child first, so the left child of node i is always node i+1 and only
// the right child index is stored.  Leaf <=> feature == -2 (sklearn's _TREE_UNDEFINED).
struct __align__(16) F16Node {
    double thr;      // float64 split threshold (sklearn Node.threshold)
    int32_t feature; // split feature or -2
    int32_t right;   // index of right child (left child = self + 1)
    int32_t c0;      // weighted class-0 sum of the node (integer valued: bootstrap counts)
    int32_t c1;      // weighted class-1 sum
    int32_t n;       // n_node_samples (distinct rows)
    int32_t depth;
};

struct __align__(16) F16StackRec {
    int32_t start, end, parent, c0, c1;
    uint32_t const_mask;  // bit f set <=> feature f is a known constant for this node
    int16_t n_const;
    uint8_t is_left;
    uint8_t pad;
    int32_t depth;
};

struct F16FitParams {
    const float* X;             // [n][dp] row-major float32 (dp = 8 or 16, zero padded)
    const uint8_t* y;           // [n] class index 0/1
    const int32_t* sorted_idx;  // [d][n] per-column argsort (best splitter only)
    const uint8_t* boot_w;      // [n_trees][n] bootstrap counts or nullptr (all weights 1)
    const uint32_t* rand_r_state;  // [n_trees]
    uint32_t* buf;              // random: [n_trees][2][n]; best: [n_trees][2][d][n]
    F16StackRec* stack;         // [n_trees][stack_cap]
    F16Node* nodes;             // [n_trees][node_cap]
    int32_t* node_count;        // [n_trees]
    uint32_t* side_global;      // [n_trees][ceil(n/32)] or nullptr (side bits live in smem)
    uint32_t* lid;              // best: [n_trees][n] global row id -> local id of the relocated node
    int32_t* err;               // device [2]: error flag, max node count over the trees (atomicMax)
    int n, d, dp, n_trees, max_features, stack_cap, node_cap, side_words;
    int salt;                   // per-launch value mixed into the leader-warp choice (f16_leader_warp)
};

struct f16_forest {
    int kind;          // 0 DT, 1 RF, 2 ET
    int n_trees;
    int d, dp;
    int node_cap;
    int64_t n_train;
    F16Node* nodes;      // device [n_trees][node_cap]
    int32_t* node_count; // device [n_trees]
    int32_t* err;        // device [2]: status code, largest node count of the trees
    int max_nodes;       // host copy of err[1], filled by f16_forest_status
    cudaEvent_t ev0, ev1;   // around the tree-building kernel when profiling is on
    int has_ev;
};

#define F16_KIND_DT 0
#define F16_KIND_RF 1
#define F16_KIND_ET 2

#define F16_STACK_CAP 4096
#define F16_SIDE_SMEM_MAX_WORDS 16384   // 64 KiB of side bits in shared memory: n <= 524288
