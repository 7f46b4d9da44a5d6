// SHARED regime of the ExtraTrees builder: one warp grows a whole subtree out of shared memory.
//
// The per-node dependency chain is what bounds a tree (its nodes are sequential), so every step
// is laid out to be short rather than wide:
//   * min/max of ALL features in one sweep: lane = (feature, row group), so the scalar feature
//     draw afterwards only shuffles constants out of registers (no data pass per drawn feature);
//   * the <= 4 candidate thresholds are counted in one sweep (lane = row, 5 ballots per 32 rows);
//   * the float64 Gini proxies of the candidates and the three Gini terms of the improvement test
//     are evaluated on different lanes at once (same operations, same rounding as the CPU's
//     sequential evaluation - only the schedule differs);
//   * stack, feature permutation and node counter stay in shared memory / registers.
#pragma once
#include "f16_tree_dev.cuh"

template <int DP, int S, class STK>
__device__ void subtree_warp_v2(Ctl& c, DrawState& ds, const STK& stk, const F16FitParams& P, F16Node* nodes,
                                const float* s_col, uint16_t (*s_idx)[S], const uint8_t* s_y) {
    constexpr int SP = S + 1;
    constexpr int G = 32 / DP;               // row groups of the all-feature min/max sweep
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const double W_total = (double)P.n;
    int sp = c.sp, node_count = c.node_count;
    uint32_t rng = ds.rng;
    const int d = P.d, max_features = P.max_features;
    const int mf = lane % DP, mg = lane / DP;
    // The splitter's Fisher-Yates permutation and its constant-feature list (16 entries of 4 bits
    // each) live in two registers for the whole subtree - every lane holds the same words - so a
    // feature draw is pure ALU work: no shared-memory round trip, no __syncwarp per swap.
    auto perm_get = [](unsigned long long w, int i) -> uint32_t { return (uint32_t)(w >> (4 * i)) & 15u; };
    auto perm_set = [](unsigned long long w, int i, uint32_t v) -> unsigned long long {
        return (w & ~(15ull << (4 * i))) | ((unsigned long long)v << (4 * i));
    };
    unsigned long long F = 0, C = 0;
#pragma unroll
    for (int i = 0; i < F16_MAX_D; i++) {
        F |= (unsigned long long)(ds.features[i] & 15) << (4 * i);
        C |= (unsigned long long)(ds.const_feats[i] & 15) << (4 * i);
    }

    while (sp > 0) {
        F16StackRec r = stk.get(sp - 1);
        __syncwarp();                           // every lane holds the record before lane 0 reuses its slot
        if (!r.pad) break;                      // back to a node that lives in global memory
        sp--;
        const int start = r.start, nn = r.end - r.start;
        const int par = r.depth & 1;
        const uint16_t* idx = s_idx[par] + start;
        const int t0 = r.c0, t1 = r.c1;
        bool split = false;
        int best_f = -2, n_left = 0, bl0 = 0, bl1 = 0;
        double best_thr = -2.0;
        int n_total = r.n_const;
        uint32_t cmask = r.const_mask;

        if (!leaf_pretest(nn, t0, t1)) {
            // ---- (a) min / max of every feature: lane owns feature mf, row group mg
            float mn = INFINITY, mx = -INFINITY;
            {
                const float* col = s_col + mf * SP;
                int i = mg;
                for (; i + 3 * G < nn; i += 4 * G) {
                    float v0 = col[idx[i]], v1 = col[idx[i + G]], v2 = col[idx[i + 2 * G]], v3 = col[idx[i + 3 * G]];
                    mn = fminf(fminf(mn, v0), fminf(v1, fminf(v2, v3)));
                    mx = fmaxf(fmaxf(mx, v0), fmaxf(v1, fmaxf(v2, v3)));
                }
                for (; i < nn; i += G) { float v = col[idx[i]]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
                for (int off = DP; off < 32; off <<= 1) {
                    mn = fminf(mn, __shfl_xor_sync(F16_FULL, mn, off));
                    mx = fmaxf(mx, __shfl_xor_sync(F16_FULL, mx, off));
                }
            }
            // ---- (b) feature draw (every lane runs the scalar loop on its own copy of the
            //      state; lane 0 applies the swaps to the shared permutation), candidates are
            //      evaluated four at a time
            int f_i = d, n_visited = 0, n_found = 0, n_drawn = 0, ncand = 0;
            const int n_known = r.n_const;
            int cf[4] = {0, 0, 0, 0};
            double ct[4] = {0.0, 0.0, 0.0, 0.0};
            float ctf[4] = {0.f, 0.f, 0.f, 0.f};     // thresholds rounded down: x <= ct  <=>  x <= ctf for float32 x
            float clo[4] = {0.f, 0.f, 0.f, 0.f}, chi[4] = {0.f, 0.f, 0.f, 0.f};
            double cr[4] = {0.0, 0.0, 0.0, 0.0};
            double best = -INFINITY;

            // rand_uniform's arithmetic (sklearn/tree/_utils.pyx:57-61), candidate j on lane j:
            // ((high - low) * r / RAND_R_MAX) + low ; threshold == max -> min (_splitter.pyx:652-653)
            auto finish_thresholds = [&]() {
                const int j = lane & 3;
                float lo = clo[0], hi = chi[0]; double r = cr[0];
#pragma unroll
                for (int m = 1; m < 4; m++) if (j == m) { lo = clo[m]; hi = chi[m]; r = cr[m]; }
                double thr = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn((double)hi, (double)lo), r), 2147483647.0), (double)lo);
                if (thr == (double)hi) thr = (double)lo;
#pragma unroll
                for (int m = 0; m < 4; m++) { ct[m] = __shfl_sync(F16_FULL, thr, m); ctf[m] = __double2float_rd(ct[m]); }
            };

            auto eval_chunk = [&](int cnt) {
                int nl[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0};
                for (int base = 0; base < nn; base += 32) {
                    const int i = base + lane;
                    const bool valid = i < nn;
                    const int li = valid ? idx[i] : 0;
                    const unsigned ym = __ballot_sync(F16_FULL, valid && s_y[li]);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k < cnt) {
                            bool left = valid && (s_col[cf[k] * SP + li] <= ctf[k]);
                            unsigned bal = __ballot_sync(F16_FULL, left);
                            nl[k] += __popc(bal); l1[k] += __popc(bal & ym);
                        }
                    }
                }
                // proxies: lane 2k -> (-w_R * g_R), lane 2k+1 -> (w_L * g_L) of candidate k
                const int k = (lane >> 1) & 3;
                int mnl = nl[0], ml1 = l1[0];
#pragma unroll
                for (int j = 1; j < 4; j++) if (k == j) { mnl = nl[j]; ml1 = l1[j]; }
                double a, b, part;
                if (lane & 1) { a = (double)(mnl - ml1); b = (double)ml1; double w = a + b; part = w * gini_of(a, b, w); }
                else { a = (double)(t0 - (mnl - ml1)); b = (double)(t1 - ml1); double w = a + b; part = (-w) * gini_of(a, b, w); }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j < cnt) {
                        double proxy = __shfl_sync(F16_FULL, part, 2 * j) - __shfl_sync(F16_FULL, part, 2 * j + 1);
                        if (proxy > best) {
                            best = proxy; best_f = cf[j]; best_thr = ct[j]; n_left = nl[j]; bl1 = l1[j]; bl0 = nl[j] - l1[j];
                        }
                    }
                }
            };

            while (f_i > n_total && (n_visited < max_features || n_visited <= n_found + n_drawn)) {
                n_visited++;
                int f_j = f16_rand_int_small(n_drawn, f_i - n_found, &rng);
                if (f_j < n_known) {
                    const uint32_t a = perm_get(F, n_drawn), b = perm_get(F, f_j);
                    F = perm_set(perm_set(F, n_drawn, b), f_j, a);
                    n_drawn++;
                    continue;
                }
                f_j += n_found;
                const int f = (int)perm_get(F, f_j);
                const float fmn = __shfl_sync(F16_FULL, mn, f), fmx = __shfl_sync(F16_FULL, mx, f);
                if (fmx <= __fadd_rn(fmn, 1e-7f)) {
                    const uint32_t b = perm_get(F, n_total);
                    F = perm_set(perm_set(F, f_j, b), n_total, (uint32_t)f);
                    n_found++; n_total++;
                    continue;
                }
                f_i--;
                {
                    const uint32_t b = perm_get(F, f_i);
                    F = perm_set(perm_set(F, f_i, (uint32_t)f), f_j, b);
                }
                // the draw consumes the generator now; the float64 division of rand_uniform is
                // deferred so that the (up to 4) thresholds of a chunk are divided on 4 lanes at once
                const double rr = (double)f16_rand_r(&rng);
#pragma unroll
                for (int j = 0; j < 4; j++) if (j == ncand) { cf[j] = f; clo[j] = fmn; chi[j] = fmx; cr[j] = rr; }
                ncand++;
                if (ncand == 4) { finish_thresholds(); eval_chunk(4); ncand = 0; }
            }
            if (ncand > 0) { finish_thresholds(); eval_chunk(ncand); }
            // features[0 : n_known] = const_feats[0 : n_known];  const_feats[n_known : n_total] =
            // features[n_known : n_total]   (sklearn/tree/_splitter.pyx:473-478), on the packed words
            {
                const unsigned long long lowm = (n_known >= 16) ? ~0ull : ((1ull << (4 * n_known)) - 1ull);
                const unsigned long long uppm = (n_total >= 16) ? ~0ull : ((1ull << (4 * n_total)) - 1ull);
                F = (F & ~lowm) | (C & lowm);
                C = (C & ~(uppm & ~lowm)) | (F & (uppm & ~lowm));
            }
            for (int i = n_known; i < n_total; i++) cmask |= 1u << perm_get(C, i);

            // ---- (c) improvement test: lane 0 parent term, lane 1 right term, lane 2 left term.
            //      For a node holding at most 1/16 of the training weight the test cannot fail: the
            //      computed (imp - B - C) is its non-negative true value minus at most 8 roundings of
            //      quantities <= 1 (> -9e-16), scaled by w_node / W <= 1/16 that is > -6e-17, four
            //      times smaller than EPSILON - so the three divisions are skipped.
            if (best_f >= 0 && improvement_certain(t0, t1, P.n)) {
                split = true;
            } else if (best_f >= 0) {
                const int role = lane % 3;
                double a, b, num, den;
                const double wn = (double)t0 + (double)t1;
                if (role == 0) { a = (double)t0; b = (double)t1; num = wn; den = W_total; }
                else if (role == 1) { a = (double)(t0 - bl0); b = (double)(t1 - bl1); num = a + b; den = wn; }
                else { a = (double)bl0; b = (double)bl1; num = a + b; den = wn; }
                const double g = gini_of(a, b, a + b);
                const double ratio = num / den;
                const double prod = ratio * g;
                const double imp = __shfl_sync(F16_FULL, g, 0), A = __shfl_sync(F16_FULL, ratio, 0);
                const double B = __shfl_sync(F16_FULL, prod, 1), C = __shfl_sync(F16_FULL, prod, 2);
                const double improvement = A * (imp - B - C);
                split = !(improvement + F16_EPS < 0.0);
            }
        }

        // ---- node record
        const int id = node_count++;
        if (id >= P.node_cap || sp + 2 > P.stack_cap) { if (lane == 0) { atomicExch(P.err, F16_ERR_OVERFLOW); c.abort = 1; } break; }
        if (lane == 0) {
            F16Node nd;
            nd.thr = split ? best_thr : -2.0;
            nd.feature = split ? best_f : -2;
            nd.right = -1;
            nd.c0 = t0; nd.c1 = t1; nd.n = nn; nd.depth = r.depth;
            nodes[id] = nd;
            if (r.parent >= 0 && !r.is_left) nodes[r.parent].right = id;
        }
        if (split) {
            // ---- stable partition idx[par] -> idx[par ^ 1]
            uint16_t* out = s_idx[par ^ 1] + start;
            const float* col = s_col + best_f * SP;
            const float best_thr_f = __double2float_rd(best_thr);
            int run_l = 0;
            for (int base = 0; base < nn; base += 32) {
                int i = base + lane;
                bool valid = i < nn;
                int li = valid ? idx[i] : 0;
                bool left = valid && (col[li] <= best_thr_f);
                unsigned bal = __ballot_sync(F16_FULL, left);
                int lrank = __popc(bal & lt);
                if (valid) {
                    if (left) out[run_l + lrank] = (uint16_t)li;
                    else out[n_left + (base - run_l) + (lane - lrank)] = (uint16_t)li;
                }
                run_l += __popc(bal);
            }
            if (lane == 0) {
                F16StackRec q;
                q.parent = id; q.depth = r.depth + 1; q.n_const = (int16_t)n_total; q.const_mask = cmask; q.pad = 1;
                q.start = start + n_left; q.end = r.end; q.c0 = t0 - bl0; q.c1 = t1 - bl1; q.is_left = 0;
                stk.put(sp, q);
                q.start = start; q.end = start + n_left; q.c0 = bl0; q.c1 = bl1; q.is_left = 1;
                stk.put(sp + 1, q);
            }
            sp += 2;
        }
        __syncwarp();
    }
    if (lane == 0) {
        c.sp = sp; c.node_count = node_count; ds.rng = rng;
#pragma unroll
        for (int i = 0; i < F16_MAX_D; i++) { ds.features[i] = (int)perm_get(F, i); ds.const_feats[i] = (int)perm_get(C, i); }
    }
}
