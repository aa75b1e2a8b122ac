/*
 * azalea_oracle.c -- CPU oracle for the Azalea self-play search hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see azalea_oracle.h).  Plain-C restatement of the
 * reference algorithm; every function cites the reference file:line it
 * follows.  Compile with -ffp-contract=off: the reference's NumPy float32
 * arithmetic rounds once per operation and never fuses multiply-add.
 */
#include "azalea_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ================================================================== hex == */

void ohex_init(ohex *g, int n)
{
    /* hex.py:145-149 */
    g->n = n;
    g->color = 1;
    g->winner = 0;
    memset(g->board, 0, sizeof(g->board));
}

int ohex_legal_moves(const ohex *g, int32_t *out)
{
    /* hex.py:151-159: empty tiles ascending, 1-based; none once won */
    int k = 0;
    if (g->winner)
        return 0;
    for (int i = 0; i < g->n * g->n; i++)
        if (g->board[i] == 0)
            out[k++] = i + 1;
    return k;
}

int ohex_result(const ohex *g)
{
    /* hex.py:161-170 */
    if (g->winner)
        return g->winner == 2 ? 1 : 3;
    return 0;
}

int ohex_neighbors(int tile, int n, int32_t out[6])
{
    /* hex.py:182-201: fixed neighbour order */
    static const int di[6] = {-1, -1, 0, 0, 1, 1};
    static const int dj[6] = {0, 1, -1, 1, -1, 0};
    int ti = tile / n, tj = tile % n, k = 0;
    for (int q = 0; q < 6; q++) {
        int ni = ti + di[q], nj = tj + dj[q];
        if (ni >= 0 && ni < n && nj >= 0 && nj < n)
            out[k++] = ni * n + nj;
    }
    return k;
}

int ohex_check_win(const int32_t *board, int n, int tile)
{
    /* hex.py:204-231: depth-first flood fill of the last move's component */
    int32_t color = board[tile];
    uint8_t connected[19 * 19];
    int32_t border[19 * 19 * 6 + 1];
    int nb = 0, imin = 9999, imax = -9999;
    if (!color)
        return -1;
    memset(connected, 0, sizeof(connected));
    border[nb++] = tile;
    while (nb) {
        int32_t neig[6];
        int t = border[--nb];
        int i, kn;
        connected[t] = 1;
        i = color == 1 ? t / n : t % n;
        if (i < imin) imin = i;
        if (i > imax) imax = i;
        if (imin == 0 && imax == n - 1)
            return color;
        kn = ohex_neighbors(t, n, neig);
        for (int q = 0; q < kn; q++) {
            if (connected[neig[q]])
                continue;
            if (board[neig[q]] == color)
                border[nb++] = neig[q];
        }
    }
    return 0;
}

int ohex_step(ohex *g, int move)
{
    /* hex.py:172-179 */
    int tile = move - 1;
    if (tile < 0 || tile >= g->n * g->n || g->board[tile] != 0 || g->winner)
        return -1;
    g->board[tile] = g->color;
    g->color = 3 - g->color;
    g->winner = ohex_check_win(g->board, g->n, tile);
    return 0;
}

void ohex_flip_board(const int32_t *board, int n, int32_t *out)
{
    /* hex.py:72-87: swap colours, mirror along the anti-diagonal:
     * out[i][j] = swap(in[n-1-j][n-1-i]) */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            int32_t v = board[(n - 1 - j) * n + (n - 1 - i)];
            out[i * n + j] = v > 0 ? 3 - v : 0;
        }
}

void ohex_flip_moves(const int32_t *moves, int k, int n, int32_t *out)
{
    /* hex.py:105-121: tile (r,c) -> (n-1-c, n-1-r); list order kept */
    for (int q = 0; q < k; q++) {
        if (moves[q] > 0) {
            int t = moves[q] - 1, r = t / n, c = t % n;
            out[q] = (n - 1 - c) * n + (n - 1 - r) + 1;
        } else {
            out[q] = 0;
        }
    }
}

/* ================================================================ stubs == */

static uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

uint32_t ostub_board_hash(const int32_t *board_view, int n)
{
    /* order-independent (sum of per-stone hashes) so a warp can reduce it */
    uint32_t s = 0;
    for (int i = 0; i < n * n; i++) {
        int32_t v = board_view[i];
        if (v)
            s += fmix32((uint32_t)(2 * i + v) * 0x9E3779B1u);
    }
    return fmix32(s ^ (uint32_t)n);
}

void ostub_eval(int mode, const int32_t *board_view, int n,
                const int32_t *moves_view, int k, float *value, float *prior)
{
    uint32_t h0 = ostub_board_hash(board_view, n);
    uint32_t wsum = 0;
    uint32_t w[19 * 19];
    if (mode == OSTUB_UNIFORM)
        *value = 0.0f;
    else if (mode == OSTUB_DYADIC)
        *value = (float)((int)((h0 >> 8) % 17u) - 8) / 8.0f;
    else
        *value = (float)(h0 >> 8) * (1.0f / 8388608.0f) - 1.0f;
    for (int j = 0; j < k; j++) {
        uint32_t t = (uint32_t)(moves_view[j] - 1);
        uint32_t hj = fmix32(h0 ^ (t * 0x9E3779B1u + 0x7F4A7C15u));
        if (mode == OSTUB_UNIFORM)
            w[j] = 1;
        else if (mode == OSTUB_DYADIC)
            w[j] = 1 + (hj >> 28);
        else
            w[j] = 1 + (hj >> 24);
        wsum += w[j];
    }
    for (int j = 0; j < k; j++)
        prior[j] = (float)w[j] / (float)wsum;
}

/* ================================================================= tree == */

otree *otree_new(int64_t max_nodes)
{
    /* search_tree.py:43-57 (the -1 fill is irrelevant: every slot is
     * written by create_child_nodes before it is read) */
    otree *t = (otree *)calloc(1, sizeof(otree));
    if (!t)
        return NULL;
    t->max_nodes = max_nodes;
    t->parent = (int32_t *)malloc(sizeof(int32_t) * max_nodes);
    t->first_child = (int32_t *)malloc(sizeof(int32_t) * max_nodes);
    t->num_children = (int32_t *)malloc(sizeof(int32_t) * max_nodes);
    t->num_visits = (float *)malloc(sizeof(float) * max_nodes);
    t->total_value = (float *)malloc(sizeof(float) * max_nodes);
    t->prior_prob = (float *)malloc(sizeof(float) * max_nodes);
    if (!t->parent || !t->first_child || !t->num_children || !t->num_visits ||
        !t->total_value || !t->prior_prob) {
        otree_free(t);
        return NULL;
    }
    otree_reset(t);
    return t;
}

void otree_free(otree *t)
{
    if (!t)
        return;
    free(t->parent);
    free(t->first_child);
    free(t->num_children);
    free(t->num_visits);
    free(t->total_value);
    free(t->prior_prob);
    free(t);
}

void otree_reset(otree *t)
{
    /* search_tree.py:59-71 */
    t->num_nodes = 1;
    t->root_id = 0;
    t->parent[0] = -1;
    t->first_child[0] = -1;
    t->num_children[0] = -1;
    t->num_visits[0] = 0;
    t->total_value[0] = 0;
    t->prior_prob[0] = 1.0f;
}

int otree_move(otree *t, int move_id)
{
    /* search_tree.py:115-132 */
    int32_t node;
    if (t->num_children[t->root_id] < 0) {
        otree_reset(t);
        return 1;
    }
    if (move_id < 0 || move_id >= t->num_children[t->root_id])
        return -1;
    node = t->first_child[t->root_id] + move_id;
    if (t->num_children[node] < 0) {
        otree_reset(t);
        return 1;
    }
    t->root_id = node;
    return 0;
}

int otree_root_stats(const otree *t, float *visits, float *total_value,
                     float *prior)
{
    /* search_tree.py:192-204 */
    int32_t k = t->num_children[t->root_id];
    int32_t fc = t->first_child[t->root_id];
    if (k < 0)
        return -1;
    for (int j = 0; j < k; j++) {
        if (visits) visits[j] = t->num_visits[fc + j];
        if (total_value) total_value[j] = t->total_value[fc + j];
        if (prior) prior[j] = t->prior_prob[fc + j];
    }
    return k;
}

static int create_child_nodes(otree *t, int32_t id, int m, const float *prior)
{
    /* search_tree.py:254-274 */
    int64_t first;
    if (t->num_nodes + m > t->max_nodes)
        return -1; /* SearchTreeFull */
    first = t->num_nodes;
    t->num_nodes += m;
    t->first_child[id] = (int32_t)first;
    t->num_children[id] = m;
    for (int j = 0; j < m; j++) {
        int64_t c = first + j;
        t->parent[c] = id;
        t->first_child[c] = -1;
        t->num_children[c] = -1;
        t->num_visits[c] = 0;
        t->total_value[c] = 0;
        t->prior_prob[c] = prior[j];
    }
    return 0;
}

/* ================================================================= mcts == */

void omcts_score_actions(const float *num_visits, const float *neg_total_value,
                         const float *prior, int k, float coef, float *score)
{
    /* mcts.py:132-135, float32, one rounding per operation:
     *   visit_gap = sqrt(sum(N)) / (1 + N)
     *   U = (coef * P) * visit_gap
     *   Q = (-W) / max(N, 1)
     *   score = Q + U
     * sum(N) is a sum of small integers, exact in any order. */
    double s = 0;
    float sq;
    for (int j = 0; j < k; j++)
        s += num_visits[j];
    sq = sqrtf((float)s);
    for (int j = 0; j < k; j++) {
        float n = num_visits[j];
        float gap = sq / (1.0f + n);
        float cp = coef * prior[j];
        float u = cp * gap;
        float q = neg_total_value[j] / (n < 1.0f ? 1.0f : n);
        score[j] = q + u;
    }
}

static void apply_virtual_loss(otree *t, int32_t node, float amount)
{
    /* mcts.py:79-92: leaf .. child-of-root; the root is not touched */
    while (node != t->root_id) {
        t->num_visits[node] += amount;
        t->total_value[node] += amount;
        node = t->parent[node];
    }
}

static void leaf_state(const ohex *g, oleaves *out, int slot)
{
    /* GameState at the leaf (hex.py:55-60) turned into the network's view
     * (mcts.py:176-181) */
    int32_t legal[19 * 19];
    int nn = g->n * g->n;
    int k = ohex_legal_moves(g, legal);
    out->color[slot] = g->color - 1;
    out->result[slot] = ohex_result(g);
    out->num_moves[slot] = k;
    if (g->color - 1 == 1) {
        ohex_flip_board(g->board, g->n, out->board_view[slot]);
        ohex_flip_moves(legal, k, g->n, out->moves_view[slot]);
    } else {
        memcpy(out->board_view[slot], g->board, sizeof(int32_t) * nn);
        memcpy(out->moves_view[slot], legal, sizeof(int32_t) * k);
    }
}

int omcts_select_batch(otree *t, const ohex *game, int batch_size, float coef,
                       oleaves *out)
{
    /* mcts.py:46-76 */
    int32_t leaf_nodes[OMAX_BATCH];
    static __thread ohex leaf_games[OMAX_BATCH];
    float nv[19 * 19], ntv[19 * 19], pr[19 * 19], score[19 * 19];
    if (batch_size > OMAX_BATCH || t->num_children[t->root_id] <= 0)
        return -1;
    out->sum_children = 0;
    out->sum_depth = 0;
    for (int i = 0; i < batch_size; i++) {
        ohex g = *game;                 /* snapshot, search_tree.py:150 */
        int32_t node = t->root_id;
        /* select_leaf, mcts.py:95-116 */
        while (t->num_children[node] > 0) {
            int32_t k = t->num_children[node], fc = t->first_child[node];
            int32_t legal[19 * 19];
            int best = 0;
            for (int j = 0; j < k; j++) {
                nv[j] = t->num_visits[fc + j];
                ntv[j] = -t->total_value[fc + j];
                pr[j] = t->prior_prob[fc + j];
            }
            omcts_score_actions(nv, ntv, pr, k, coef, score);
            for (int j = 1; j < k; j++)  /* np.argmax: first maximum */
                if (score[j] > score[best])
                    best = j;
            /* ForwardSearchIterator.step, search_tree.py:298-308 */
            if (ohex_legal_moves(&g, legal) != k)
                return -2;
            ohex_step(&g, legal[best]);
            node = fc + best;
            out->sum_children += k;
            out->sum_depth += 1;
        }
        apply_virtual_loss(t, node, 1.0f);      /* mcts.py:69 */
        leaf_nodes[i] = node;
        leaf_games[i] = g;
    }
    for (int i = 0; i < batch_size; i++)        /* mcts.py:72 */
        apply_virtual_loss(t, leaf_nodes[i], -1.0f);
    /* deduplicate_leaves, mcts.py:139-152 */
    out->count = 0;
    for (int i = 0; i < batch_size; i++) {
        int seen = 0;
        for (int q = 0; q < out->count; q++)
            if (out->node[q] == leaf_nodes[i])
                seen = 1;
        if (seen)
            continue;
        out->node[out->count] = leaf_nodes[i];
        leaf_state(&leaf_games[i], out, out->count);
        out->count++;
    }
    return out->count;
}

int omcts_expand_backup(otree *t, const oleaves *lv, const float *value,
                        const float *prior, int stride, float *sum_values)
{
    float vals[OMAX_BATCH];
    float sv = 0;
    /* evaluate_batch, mcts.py:192-200: terminal rows are worth -1 */
    for (int i = 0; i < lv->count; i++)
        vals[i] = lv->result[i] != 0 ? -1.0f : value[i];
    /* expand_batch, mcts.py:226-239 */
    for (int i = 0; i < lv->count; i++) {
        int32_t node = lv->node[i];
        if (t->num_children[node] > 0)
            return -2; /* expanded node is not leaf */
        if (t->num_children[node] != 0)
            if (create_child_nodes(t, node, lv->num_moves[i],
                                   prior + (size_t)i * stride))
                return -1;
    }
    /* backup_batch, mcts.py:242-255: leaf .. root inclusive */
    for (int i = 0; i < lv->count; i++) {
        int32_t node = lv->node[i];
        float v = vals[i];
        for (;;) {
            t->total_value[node] += v;
            t->num_visits[node] += 1.0f;
            v = -v;
            if (node == t->root_id)
                break;
            node = t->parent[node];
        }
        sv += vals[i];
    }
    if (sum_values)
        *sum_values = sv;
    return 0;
}

int omcts_root_leaf(const otree *t, const ohex *game, oleaves *out)
{
    /* evaluate_root, mcts.py:18-27 */
    out->count = 1;
    out->node[0] = t->root_id;
    out->sum_children = 0;
    out->sum_depth = 0;
    leaf_state(game, out, 0);
    return 1;
}

int omcts_expand_root(otree *t, const oleaves *lv, const float *prior)
{
    /* mcts.py:25-26: len(prior_prob[0]) children, value discarded */
    return create_child_nodes(t, t->root_id, lv->num_moves[0], prior);
}

static void eval_stub(const oleaves *lv, int n, int mode, float *value,
                      float *prior, int stride)
{
    for (int i = 0; i < lv->count; i++) {
        if (lv->result[i] != 0) {
            value[i] = -1.0f;
            continue;
        }
        ostub_eval(mode, lv->board_view[i], n, lv->moves_view[i],
                   lv->num_moves[i], &value[i], prior + (size_t)i * stride);
    }
}

typedef struct ocounters {
    int64_t sum_children, sum_depth, unique_leaves;
} ocounters;

static int sample_paths_stub(otree *t, const ohex *game, int num_simulations,
                             int batch_size, float coef, int stub_mode,
                             float *search_value, oleaves *lv, float *prior,
                             ocounters *cnt)
{
    /* mcts.py:258-293 */
    int num_batches = num_simulations / batch_size + 1;
    int stride = 19 * 19;
    float value[OMAX_BATCH];
    float sv_total = 0;
    if (t->num_children[t->root_id] < 0) {
        omcts_root_leaf(t, game, lv);
        eval_stub(lv, game->n, stub_mode, value, prior, stride);
        if (omcts_expand_root(t, lv, prior))
            return -1;
    }
    for (int b = 0; b < num_batches; b++) {
        float sv;
        int rc = omcts_select_batch(t, game, batch_size, coef, lv);
        if (rc < 0)
            return rc;
        eval_stub(lv, game->n, stub_mode, value, prior, stride);
        rc = omcts_expand_backup(t, lv, value, prior, stride, &sv);
        if (rc < 0)
            return rc;
        sv_total += sv;
        if (cnt) {
            cnt->sum_children += lv->sum_children;
            cnt->sum_depth += lv->sum_depth;
            cnt->unique_leaves += lv->count;
        }
    }
    if (search_value)
        *search_value = sv_total / (float)(num_batches * batch_size);
    return 0;
}

int omcts_sample_paths_stub(otree *t, const ohex *game, int num_simulations,
                            int batch_size, float coef, int stub_mode,
                            float *search_value)
{
    oleaves *lv = (oleaves *)malloc(sizeof(oleaves));
    float *prior = (float *)malloc(sizeof(float) * OMAX_BATCH * 19 * 19);
    int rc = -1;
    if (lv && prior)
        rc = sample_paths_stub(t, game, num_simulations, batch_size, coef,
                               stub_mode, search_value, lv, prior, NULL);
    free(lv);
    free(prior);
    return rc;
}

/* ========================================================= CPU baseline == */

typedef struct obench_job {
    int n, num_simulations, batch_size, stub_mode, exploration_depth;
    float coef;
    int64_t max_nodes;
    uint64_t seed;
    int first_game, num_games;
    int64_t plies, sims;
    ocounters cnt;
    int failed;
} obench_job;

static uint32_t lcg_next(uint64_t *s)
{
    *s = *s * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(*s >> 33);
}

static void *obench_worker(void *arg)
{
    obench_job *job = (obench_job *)arg;
    oleaves *lv = (oleaves *)malloc(sizeof(oleaves));
    float *prior = (float *)malloc(sizeof(float) * OMAX_BATCH * 19 * 19);
    otree *t = otree_new(job->max_nodes);
    float visits[19 * 19];
    int32_t legal[19 * 19];
    if (!lv || !prior || !t) {
        job->failed = 1;
        goto done;
    }
    for (int gi = 0; gi < job->num_games; gi++) {
        /* play_game, play_game.py:44-54, one agent on both sides
         * (policy_trainer.py:72-75): one tree reused across plies */
        uint64_t rng = job->seed * 0x9E3779B97F4A7C15ull +
                       (uint64_t)(job->first_game + gi) * 2654435761ull + 1;
        ohex g;
        int sims_per_move =
            (job->num_simulations / job->batch_size + 1) * job->batch_size;
        ohex_init(&g, job->n);
        otree_reset(t);             /* Policy.reset, policy.py:72-76 */
        for (int ply = 0; ply < 300 && !g.winner; ply++) {
            int k, move_id = 0;
            double total = 0;
            if (sample_paths_stub(t, &g, job->num_simulations,
                                  job->batch_size, job->coef, job->stub_mode,
                                  NULL, lv, prior, &job->cnt) < 0) {
                /* SearchTreeFull: game dropped, parallel_player.py:71-76 */
                break;
            }
            job->sims += sims_per_move;
            k = otree_root_stats(t, visits, NULL, NULL);
            for (int j = 0; j < k; j++)
                total += visits[j];
            if (ply < job->exploration_depth) {
                /* temperature 1: pi proportional to visits */
                double r = (lcg_next(&rng) / 2147483648.0) * total, acc = 0;
                for (int j = 0; j < k; j++) {
                    acc += visits[j];
                    if (r < acc) {
                        move_id = j;
                        break;
                    }
                }
            } else {
                /* temperature 0: uniform over the arg-max ties */
                float best = -1;
                int ties = 0, pick;
                for (int j = 0; j < k; j++)
                    if (visits[j] > best) {
                        best = visits[j];
                        ties = 1;
                    } else if (visits[j] == best) {
                        ties++;
                    }
                pick = (int)(lcg_next(&rng) % (uint32_t)ties);
                for (int j = 0; j < k; j++)
                    if (visits[j] == best && pick-- == 0) {
                        move_id = j;
                        break;
                    }
            }
            ohex_legal_moves(&g, legal);
            otree_move(t, move_id);     /* policy.py:170-176 */
            ohex_step(&g, legal[move_id]);
            job->plies++;
        }
    }
done:
    otree_free(t);
    free(lv);
    free(prior);
    return NULL;
}

int64_t obench_selfplay_stub(int n, int num_games, int threads,
                             int num_simulations, int batch_size, float coef,
                             int stub_mode, int exploration_depth,
                             int64_t max_nodes, uint64_t seed,
                             int64_t *simulations, double *seconds,
                             int64_t *sum_children, int64_t *sum_depth,
                             int64_t *unique_leaves)
{
    pthread_t *th;
    obench_job *jobs;
    struct timespec t0, t1;
    int64_t plies = 0, sims = 0, sc = 0, sd = 0, ul = 0;
    if (threads < 1)
        threads = 1;
    if (threads > num_games)
        threads = num_games;
    th = (pthread_t *)calloc(threads, sizeof(pthread_t));
    jobs = (obench_job *)calloc(threads, sizeof(obench_job));
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < threads; i++) {
        int lo = (int)((int64_t)num_games * i / threads);
        int hi = (int)((int64_t)num_games * (i + 1) / threads);
        jobs[i].n = n;
        jobs[i].num_simulations = num_simulations;
        jobs[i].batch_size = batch_size;
        jobs[i].coef = coef;
        jobs[i].stub_mode = stub_mode;
        jobs[i].exploration_depth = exploration_depth;
        jobs[i].max_nodes = max_nodes;
        jobs[i].seed = seed;
        jobs[i].first_game = lo;
        jobs[i].num_games = hi - lo;
        pthread_create(&th[i], NULL, obench_worker, &jobs[i]);
    }
    for (int i = 0; i < threads; i++) {
        pthread_join(th[i], NULL);
        plies += jobs[i].plies;
        sims += jobs[i].sims;
        sc += jobs[i].cnt.sum_children;
        sd += jobs[i].cnt.sum_depth;
        ul += jobs[i].cnt.unique_leaves;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (simulations) *simulations = sims;
    if (seconds)
        *seconds = (double)(t1.tv_sec - t0.tv_sec) +
                   1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (sum_children) *sum_children = sc;
    if (sum_depth) *sum_depth = sd;
    if (unique_leaves) *unique_leaves = ul;
    free(th);
    free(jobs);
    return plies;
}
