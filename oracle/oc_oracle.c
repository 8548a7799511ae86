/*
 * oc_oracle.c -- TEST INFRASTRUCTURE ONLY: plain-C CPU restatement of the gym-comm environment
 * step + observation path, batched over independent envs (pthreads over envs).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this.  The
 * product library (gym_comm_b200/liboc_b200.so) shares no code with it: this file parses the
 * level text itself, runs its own BFS and keeps an unpacked per-object world.
 *
 * Parity status: PINNED.  The reference has no tests/golden vectors for this path, so this
 * restatement is pinned against outputs of the reference itself: tests/test_c_oracle.py replays
 * every trace in tests/golden (recorded from the live reference by oracle/record_golden.py)
 * bit-exactly, and cross-checks it against oracle/spec_model.py on fresh runs.
 *
 * Reference anchors (file:line relative to the reference root):
 *   level parsing        gym_cooking/envs/overcooked_environment.py:100-178
 *   step                 gym_cooking/envs/overcooked_environment.py:211-241
 *   collisions           :543-613          interact  gym_cooking/utils/interact.py:4-75
 *   done / reward        :243-270, :399-432
 *   reward shaping       :272-397          path distance  gym_cooking/utils/world.py:61-131
 *   wrapper + obs        gym_comm/envs/overcooked_env.py:105-159, 207-297
 *
 * Encodings: content bits Tomato 1, Lettuce 2, Onion 4, Plate 8; tiles 0 Floor 1 Counter
 * 2 Cutboard 3 Delivery; nav 0 (0,+1) 1 (0,-1) 2 (-1,0) 3 (+1,0).
 *
 * Device-RNG twins (so auto-reset / fused rollouts can be compared with the CUDA path at scale):
 * Philox4x32-10, key = (seed lo, seed hi).  Random placement: counter (env, episode, 'RESE', j/4),
 * draw j -> index mulhi(word, ncounters - j) among the Counter tiles not yet taken.  Rollout
 * actions: counter (env, global step, 'ACTS', 0): nav_k = bits [2k,2k+2) of word 0,
 * comm_0/1 = mulhi(word 1/2, C).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

#define MAXA 4
#define MAXO 8
#define MAXS 32
#define MAXCELL 256
#define FOODS 7
#define PLATE 8

static const int NAVX[4] = {0, 0, -1, 1}, NAVY[4] = {1, -1, 0, 0};   /* world.py:16 */

typedef struct {
    int contents, chopped, x, y, held_by /* -1 none */, alive, stamp;
} Obj;

typedef struct {
    int t, episodes;
    int ax[MAXA], ay[MAXA];
    Obj o[MAXO];
    int rank[16], nkeys, next_stamp;     /* world.objects key creation order / list order (world.py:236-237) */
    int completed[MAXS], count[MAXS];
    int comm[2];                         /* message index, -1 = all-zero vector */
    int last_completed;
} Env;

typedef struct {
    /* config */
    int N, A, T, C, comm_on, ego_led, fow;
    int can_move[MAXA], allergic[MAXA], blind[MAXA];
    uint64_t seed;
    /* level */
    int W, H, M;
    int tile[MAXCELL];
    int startx[MAXA], starty[MAXA];
    int nobj, obj_bits[MAXO], obj_x[MAXO], obj_y[MAXO]; /* x = -1: random counter at reset */
    int nrandom, random_slot[MAXO];
    int ncounters, counter_x[MAXCELL], counter_y[MAXCELL];
    int ndelivery, delivery_x[MAXCELL], delivery_y[MAXCELL];
    int nitems, items[4];
    int pd[MAXCELL][MAXCELL];
    /* subtasks */
    int S, kind[MAXS] /* 0 Chop 1 Merge 2 Deliver */, goal_c[MAXS], goal_ch[MAXS], arg0[MAXS];
    int F, off[11];
    Env* env;
    uint32_t rollout_step;
} Batch;

/* ---------------------------------------------------------------------------------- Philox */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ---------------------------------------------------------------------------------- threads
 * plain pthreads (this image's gcc has no libgomp): run fn(ctx, tid, nthreads) on every core */
typedef void (*par_fn)(void* ctx, int tid, int nth);
typedef struct { par_fn fn; void* ctx; int tid, nth; } ParArg;
static void* par_tramp(void* a) { ParArg* p = (ParArg*)a; p->fn(p->ctx, p->tid, p->nth); return NULL; }
static int host_threads(void) {
    const char* e = getenv("OCO_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
static void parallel_run(par_fn fn, void* ctx, int work_items) {
    int nth = host_threads();
    if (nth > work_items) nth = work_items > 0 ? work_items : 1;
    if (nth == 1) { fn(ctx, 0, 1); return; }
    pthread_t th[256]; ParArg arg[256];
    for (int t = 0; t < nth; ++t) { arg[t].fn = fn; arg[t].ctx = ctx; arg[t].tid = t; arg[t].nth = nth; pthread_create(&th[t], NULL, par_tramp, &arg[t]); }
    for (int t = 0; t < nth; ++t) pthread_join(th[t], NULL);
}
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* ---------------------------------------------------------------------------------- level */
static int bit_of_name(const char* s, int n) {
    if (n == 6 && !strncmp(s, "Tomato", 6)) return 1;
    if (n == 7 && !strncmp(s, "Lettuce", 7)) return 2;
    if (n == 5 && !strncmp(s, "Onion", 5)) return 4;
    if (n == 5 && !strncmp(s, "Plate", 5)) return 8;
    return 0;
}
static int bits_of_arg(const char* s, int n) {       /* "Lettuce-Plate" */
    int bits = 0, i = 0;
    while (i < n) {
        int j = i;
        while (j < n && s[j] != '-') ++j;
        bits |= bit_of_name(s + i, j - i);
        i = j + 1;
    }
    return bits;
}

/* subtasks: "Deliver(Plate-Tomato);Merge(Tomato, Plate);Chop(Tomato)"  (str() of the reference's
 * all_subtasks).  Goal template = union of the args with every Food chopped
 * (navigation_planner/utils.py:161-209). */
static int parse_subtasks(Batch* b, const char* s) {
    b->S = 0;
    while (*s) {
        const char* e = strchr(s, ';');
        int len = e ? (int)(e - s) : (int)strlen(s);
        const char* lp = memchr(s, '(', len);
        if (!lp || b->S >= MAXS) return -1;
        int k = (int)(lp - s), i = b->S;
        if (k == 4 && !strncmp(s, "Chop", 4)) b->kind[i] = 0;
        else if (k == 5 && !strncmp(s, "Merge", 5)) b->kind[i] = 1;
        else if (k == 7 && !strncmp(s, "Deliver", 7)) b->kind[i] = 2;
        else return -1;
        const char* a = lp + 1;
        const char* end = s + len - 1;              /* ')' */
        const char* comma = memchr(a, ',', end - a);
        int bits0 = bits_of_arg(a, (int)((comma ? comma : end) - a));
        int bits = bits0;
        if (comma) bits |= bits_of_arg(comma + 2, (int)(end - (comma + 2)));
        b->goal_c[i] = bits; b->goal_ch[i] = bits & FOODS; b->arg0[i] = bits0;
        b->S++;
        s += len + (e ? 1 : 0);
    }
    return b->S > 0 ? 0 : -1;
}

static int parse_level(Batch* b, const char* text) {
    int phase = 1, y = 0, w = 0, nstart = 0, r0 = -1;
    b->nobj = 0; b->nrandom = 0;
    const char* p = text;
    int nfixed = 0;
    int rnd_bits[MAXO], nr = 0;
    while (1) {
        const char* e = strchr(p, '\n');
        int len = e ? (int)(e - p) : (int)strlen(p);
        if (len == 0) { phase++; }
        else if (phase == 1) {
            for (int x = 0; x < len; ++x) {
                char ch = p[x];
                int c = y * len + x, bit = ch == 't' ? 1 : ch == 'l' ? 2 : ch == 'o' ? 4 : ch == 'p' ? 8 : 0;
                if (c >= MAXCELL) return -1;
                if (bit) {
                    b->tile[c] = 1;
                    if (b->nobj >= MAXO) return -1;
                    b->obj_bits[b->nobj] = bit; b->obj_x[b->nobj] = x; b->obj_y[b->nobj] = y; b->nobj++; nfixed++;
                } else b->tile[c] = ch == '-' ? 1 : ch == '/' ? 2 : ch == '*' ? 3 : 0;
            }
            w = len; y++;
        } else if (phase == 2) {
            if (r0 < 0) {
                if (len == 12 && !strncmp(p, "SimpleTomato", 12)) r0 = 1;
                else if (len == 13 && !strncmp(p, "SimpleLettuce", 13)) r0 = 2;
                else if (len == 5 && !strncmp(p, "Salad", 5)) r0 = 3;
                else if (len == 10 && !strncmp(p, "OnionSalad", 10)) r0 = 7;
                else return -1;
            }
        } else if (phase == 3) {
            if (nstart < b->A) { int sx, sy; if (sscanf(p, "%d %d", &sx, &sy) != 2) return -1; b->startx[nstart] = sx; b->starty[nstart] = sy; nstart++; }
        } else if (phase == 4) {
            for (int x = 0; x < len; ++x) {
                char ch = p[x];
                int bit = ch == 't' ? 1 : ch == 'l' ? 2 : ch == 'o' ? 4 : ch == 'p' ? 8 : 0;
                if (bit) rnd_bits[nr++] = bit;
            }
        }
        if (!e) break;
        p = e + 1;
    }
    (void)nfixed;
    for (int j = 0; j < nr; ++j) {
        if (b->nobj >= MAXO) return -1;
        b->random_slot[b->nrandom++] = b->nobj;
        b->obj_bits[b->nobj] = rnd_bits[j]; b->obj_x[b->nobj] = -1; b->obj_y[b->nobj] = -1; b->nobj++;
    }
    if (nstart < b->A || r0 < 0) return -1;
    b->W = w; b->H = y; b->M = 2 * (w + y) + 1;                 /* perimeter + 1 (:178, :274) */
    b->ncounters = b->ndelivery = 0;
    for (int yy = 0; yy < b->H; ++yy)
        for (int x = 0; x < b->W; ++x) {
            int t = b->tile[yy * b->W + x];
            if (t == 1) { b->counter_x[b->ncounters] = x; b->counter_y[b->ncounters] = yy; b->ncounters++; }
            if (t == 3) { b->delivery_x[b->ndelivery] = x; b->delivery_y[b->ndelivery] = yy; b->ndelivery++; }
        }
    /* shaping items: Plate + recipes[0].contents sorted by name: Lettuce < Onion < Tomato (:319-321) */
    b->nitems = 0; b->items[b->nitems++] = PLATE;
    if (r0 & 2) b->items[b->nitems++] = 2;
    if (r0 & 4) b->items[b->nitems++] = 4;
    if (r0 & 1) b->items[b->nitems++] = 1;
    /* path distances (world.py:61-131): BFS over floor; collidable dst = 1 + nearest floor neighbour;
     * non-floor src -> MAX_PATH (node missing, exception swallowed) */
    int n = b->W * b->H;
    for (int s = 0; s < n; ++s) {
        for (int d = 0; d < n; ++d) b->pd[s][d] = b->M;
        if (b->tile[s] != 0) continue;
        int dist[MAXCELL], q[MAXCELL], qh = 0, qt = 0;
        for (int i = 0; i < n; ++i) dist[i] = -1;
        dist[s] = 0; q[qt++] = s;
        while (qh < qt) {
            int u = q[qh++], ux = u % b->W, uy = u / b->W;
            for (int a = 0; a < 4; ++a) {
                int vx = ux + NAVX[a], vy = uy + NAVY[a];
                if (vx < 0 || vy < 0 || vx >= b->W || vy >= b->H) continue;
                int v = vy * b->W + vx;
                if (b->tile[v] == 0 && dist[v] < 0) { dist[v] = dist[u] + 1; q[qt++] = v; }
            }
        }
        for (int d = 0; d < n; ++d) {
            int best = b->M;
            if (b->tile[d] == 0) { if (dist[d] >= 0) best = dist[d]; }
            else {
                int dx = d % b->W, dy = d / b->W;
                for (int a = 0; a < 4; ++a) {
                    int vx = dx + NAVX[a], vy = dy + NAVY[a];
                    if (vx < 0 || vy < 0 || vx >= b->W || vy >= b->H) continue;
                    int v = vy * b->W + vx;
                    if (b->tile[v] == 0 && dist[v] >= 0 && dist[v] + 1 < best) best = dist[v] + 1;
                }
            }
            b->pd[s][d] = best < b->M ? best : b->M;
        }
    }
    return 0;
}

static int PD(const Batch* b, int x0, int y0, int x1, int y1) { return b->pd[y0 * b->W + x0][y1 * b->W + x1]; }
static int TILE(const Batch* b, int x, int y) { return (x < 0 || y < 0 || x >= b->W || y >= b->H) ? 1 : b->tile[y * b->W + x]; }

/* ---------------------------------------------------------------------------------- reset */
static void insert_obj(Env* e, Obj* o) {            /* World.insert (world.py:236-237) */
    o->stamp = ++e->next_stamp;
    if (e->rank[o->contents] == 0) e->rank[o->contents] = ++e->nkeys;
}

static void env_reset(const Batch* b, Env* e, int env_id, const int32_t* placements /* R cells or NULL */) {
    e->t = 0;
    for (int k = 0; k < b->A; ++k) { e->ax[k] = b->startx[k]; e->ay[k] = b->starty[k]; }
    memset(e->rank, 0, sizeof(e->rank));
    e->nkeys = 0; e->next_stamp = 0;
    memset(e->completed, 0, sizeof(e->completed));
    memset(e->count, 0, sizeof(e->count));
    int rx[MAXO], ry[MAXO];
    if (b->nrandom > 0) {
        if (placements) {
            for (int j = 0; j < b->nrandom; ++j) { rx[j] = placements[j] % b->W; ry[j] = placements[j] / b->W; }
        } else {
            uint32_t r[8];
            philox4x32_10((uint32_t)env_id, (uint32_t)e->episodes, 0x52455345u, 0u, (uint32_t)b->seed, (uint32_t)(b->seed >> 32), r);
            philox4x32_10((uint32_t)env_id, (uint32_t)e->episodes, 0x52455345u, 1u, (uint32_t)b->seed, (uint32_t)(b->seed >> 32), r + 4);
            int taken[MAXCELL];
            memset(taken, 0, sizeof(taken));
            for (int j = 0; j < b->nrandom; ++j) {
                uint32_t idx = mulhi(r[j], (uint32_t)(b->ncounters - j));
                int c = -1;
                for (int i = 0; i < b->ncounters; ++i) {          /* idx-th Counter not taken yet */
                    if (taken[i]) continue;
                    if (idx == 0) { c = i; break; }
                    --idx;
                }
                taken[c] = 1;
                rx[j] = b->counter_x[c]; ry[j] = b->counter_y[c];
            }
        }
    }
    int jr = 0;
    for (int s = 0; s < MAXO; ++s) e->o[s].alive = 0;
    for (int s = 0; s < b->nobj; ++s) {
        Obj* o = &e->o[s];
        o->contents = b->obj_bits[s]; o->chopped = 0; o->held_by = -1; o->alive = 1;
        if (b->obj_x[s] < 0) { o->x = rx[jr]; o->y = ry[jr]; jr++; } else { o->x = b->obj_x[s]; o->y = b->obj_y[s]; }
        insert_obj(e, o);
    }
}

/* ---------------------------------------------------------------------------------- step */
static Obj* unheld_at(Env* e, int nobj, int x, int y) {
    for (int s = 0; s < nobj; ++s)
        if (e->o[s].alive && e->o[s].held_by < 0 && e->o[s].x == x && e->o[s].y == y) return &e->o[s];
    return NULL;
}
static Obj* held_by(Env* e, int nobj, int k) {
    for (int s = 0; s < nobj; ++s)
        if (e->o[s].alive && e->o[s].held_by == k) return &e->o[s];
    return NULL;
}
static int popcount4(int v) { return (v & 1) + ((v >> 1) & 1) + ((v >> 2) & 1) + ((v >> 3) & 1); }

static int at_delivery(const Batch* b, const Env* e, int c, int ch) {      /* first Delivery tile only (:259,:402) */
    for (int s = 0; s < b->nobj; ++s) {
        const Obj* o = &e->o[s];
        if (o->alive && o->contents == c && o->chopped == ch && o->x == b->delivery_x[0] && o->y == b->delivery_y[0]) return 1;
    }
    return 0;
}
/* distinct locations of objects equal to the template: len(set(get_object_locs held + unheld)) (world.py:278-291) */
static int count_locs(const Batch* b, const Env* e, int c, int ch, int* fx, int* fy) {
    int n = 0, xs[MAXO], ys[MAXO];
    for (int s = 0; s < b->nobj; ++s) {
        const Obj* o = &e->o[s];
        if (!(o->alive && o->contents == c && o->chopped == ch)) continue;
        int dup = 0;
        for (int i = 0; i < n; ++i) if (xs[i] == o->x && ys[i] == o->y) dup = 1;
        if (!dup) { xs[n] = o->x; ys[n] = o->y; n++; }
    }
    if (n > 0 && fx) { *fx = xs[0]; *fy = ys[0]; }
    return n;
}

static double shaping(const Batch* b, const Env* e, int k) {               /* :272-397 */
    const int M = b->M, ax = e->ax[k], ay = e->ay[k];
    int have_float = 0; double tp = 0.0;      /* Python: tp is int 0 until a float is added; 0 + x == x */
    int nU = 0, minU = 1 << 30;
    for (int i = 0; i < b->S; ++i)
        if (b->kind[i] == 0 && !e->completed[i]) {
            int fx, fy;
            if (count_locs(b, e, b->arg0[i], 0, &fx, &fy) == 0) continue;  /* the reference would raise IndexError */
            int d = PD(b, ax, ay, fx, fy);
            if (d < minU) minU = d;
            nU++;
        }
    if (nU > 0) { tp += (double)((minU + M) + (nU - 1) * 2 * M) / (double)M; have_float = 1; }   /* :303-304 */
    int nP = 0, minP = 1 << 30;
    for (int i = 0; i < b->nitems; ++i)
        for (int j = i + 1; j < b->nitems; ++j) {                          /* combinations(items, 2) :340-356 */
            int n1 = 0, n2 = 0, m = M;
            for (int s = 0; s < b->nobj; ++s) {
                const Obj* o1 = &e->o[s];
                if (!o1->alive) continue;
                if (o1->contents & b->items[j]) n2++;
                if (!(o1->contents & b->items[i])) continue;
                n1++;
                for (int r = 0; r < b->nobj; ++r) {
                    const Obj* o2 = &e->o[r];
                    if (!o2->alive || !(o2->contents & b->items[j])) continue;
                    int d = PD(b, o1->x, o1->y, o2->x, o2->y);
                    if (d < m) m = d;
                }
            }
            if (n1 > 0 && n2 > 0) { if (m != 0) { if (m < minP) minP = m; nP++; } }
            else { if (M < minP) minP = M; nP++; }
        }
    if (nP > 0) {                                                          /* :359-363 */
        if (!have_float) tp += (double)(minP + (nP - 1) * M) / (double)M;
        else tp += (double)(nP * M) / (double)M;
    }
    for (int i = 0; i < b->S; ++i)                                         /* :370-395 */
        if (b->kind[i] == 2 && !e->completed[i]) {
            int dx, dy;
            if (count_locs(b, e, b->goal_c[i], b->goal_ch[i], &dx, &dy) == 0) { tp += 2.0; continue; }
            int d = PD(b, ax, ay, dx, dy) + abs(ax - dx) + abs(ay - dy);
            if (d == 0) {
                int best = 1 << 30;
                for (int t = 0; t < b->ndelivery; ++t) {
                    int v = PD(b, ax, ay, b->delivery_x[t], b->delivery_y[t]) + abs(ax - b->delivery_x[t]) + abs(ay - b->delivery_y[t]);
                    if (v < best) best = v;
                }
                tp += (double)best / (double)M;
            } else tp += (double)d / (double)M + 1.0;
        }
    return tp;
}

static void env_step(const Batch* b, Env* e, const int* nav, const int* comm, double* reward, int* done) {
    /* comm write (overcooked_env.py:227-246) */
    e->comm[0] = b->comm_on ? comm[0] : -1;
    e->comm[1] = (b->comm_on && !b->ego_led) ? comm[1] : -1;
    int actx[MAXA], acty[MAXA], has[MAXA];
    for (int k = 0; k < b->A; ++k) {                                       /* :248-262 */
        has[k] = b->can_move[k];
        actx[k] = has[k] ? NAVX[nav[k]] : 0; acty[k] = has[k] ? NAVY[nav[k]] : 0;
    }
    e->t += 1;                                                             /* :213 */
    /* collisions on the original actions (:543-613); off-grid = blocked */
    int nx[MAXA], ny[MAXA], ex[MAXA];
    for (int k = 0; k < b->A; ++k) {
        int cx = e->ax[k] + actx[k], cy = e->ay[k] + acty[k];
        if (TILE(b, cx, cy) == 0) { nx[k] = cx; ny[k] = cy; } else { nx[k] = e->ax[k]; ny[k] = e->ay[k]; }
        ex[k] = 1;
    }
    for (int i = 0; i < b->A; ++i)
        for (int j = i + 1; j < b->A; ++j) {
            if (nx[i] == nx[j] && ny[i] == ny[j]) {
                if (nx[i] == e->ax[i] && ny[i] == e->ay[i] && has[i]) ex[j] = 0;
                else if (nx[j] == e->ax[j] && ny[j] == e->ay[j] && has[j]) ex[i] = 0;
                else { ex[i] = 0; ex[j] = 0; }
            } else if (e->ax[i] == nx[j] && e->ay[i] == ny[j] && e->ax[j] == nx[i] && e->ay[j] == ny[i]) { ex[i] = 0; ex[j] = 0; }
        }
    /* interact, agents in order (interact.py:4-75) */
    for (int k = 0; k < b->A; ++k) {
        if (!has[k] || !ex[k]) continue;
        int tx = e->ax[k] + actx[k], ty = e->ay[k] + acty[k];
        if (tx < 0) tx = 0; if (ty < 0) ty = 0; if (tx >= b->W) tx = b->W - 1; if (ty >= b->H) ty = b->H - 1;   /* world.py:317-320 */
        int tt = b->tile[ty * b->W + tx];
        Obj* h = held_by(e, b->nobj, k);
        if (tt == 0) { e->ax[k] = tx; e->ay[k] = ty; if (h) { h->x = tx; h->y = ty; } continue; }
        Obj* o = unheld_at(e, b->nobj, tx, ty);
        if (h) {
            int h_done = (h->contents & FOODS) == h->chopped;
            if (tt == 3) {                                                 /* deliver (:25-30, core.py:232-237) */
                if (popcount4(h->contents) > 1 && h_done) { h->x = tx; h->y = ty; h->held_by = -1; }
            } else if (o) {                                                /* merge (:33-42, core.py:240-257) */
                if (!(h->contents & o->contents & PLATE) && h_done && (o->contents & FOODS) == o->chopped) {
                    o->alive = 0;
                    h->contents |= o->contents; h->chopped |= o->chopped;
                    insert_obj(e, h);
                }
            } else if (tt == 2 && (h->contents == 1 || h->contents == 2 || h->contents == 4) && !h->chopped) {
                h->chopped = h->contents;                                  /* chop in hand (:50-54) */
            } else { h->x = tx; h->y = ty; h->held_by = -1; }              /* put down (:55-59) */
        } else if (o && tt != 3 && !b->allergic[k]) {                      /* pick up (:64-71, agent.py:296-305) */
            o->held_by = k; o->x = e->ax[k]; o->y = e->ay[k];
        }
    }
    /* done (:243-270) */
    int d;
    if (b->T && e->t >= b->T) d = 1;
    else { d = 1; for (int i = 0; i < b->S; ++i) if (b->kind[i] == 2 && !at_delivery(b, e, b->goal_c[i], b->goal_ch[i])) { d = 0; break; } }
    /* reward (:399-432) */
    int rew = 0;
    for (int i = 0; i < b->S; ++i) {
        int r = 0;
        if (b->kind[i] == 2) { if (at_delivery(b, e, b->goal_c[i], b->goal_ch[i])) r = 3; }
        else {
            int c = count_locs(b, e, b->goal_c[i], b->goal_ch[i], NULL, NULL);
            if (c > e->count[i]) r = 1;
            e->count[i] = c;
        }
        rew += r;
        if (r) e->completed[i] = 1;
    }
    double s0 = shaping(b, e, 0), s1 = shaping(b, e, 1);
    *reward = ((double)rew - s0) - s1;                                     /* overcooked_env.py:282 */
    *done = d;
}

/* ---------------------------------------------------------------------------------- obs */
static void env_obs(const Batch* b, const Env* e, int k, double* out) {    /* overcooked_env.py:105-159 */
    for (int i = 0; i < b->F; ++i) out[i] = 0.0;
    int blind = b->blind[k];
    int dx[4] = {0, 0, 0, 0}, dy[4] = {0, 0, 0, 0}, st[4] = {0, 0, 0, 0}, hid[4] = {1, 1, 1, 1};
    if (!blind) {
        /* iterate world.objects in dict order: sort alive objects by (key rank, stamp); last writer wins */
        int idx[MAXO], n = 0;
        for (int s = 0; s < b->nobj; ++s) if (e->o[s].alive) idx[n++] = s;
        for (int i = 1; i < n; ++i) {
            int v = idx[i], j = i - 1;
            while (j >= 0) {
                const Obj *p = &e->o[idx[j]], *q = &e->o[v];
                int kp = e->rank[p->contents] * 256 + p->stamp, kq = e->rank[q->contents] * 256 + q->stamp;
                if (kp <= kq) break;
                idx[j + 1] = idx[j]; --j;
            }
            idx[j + 1] = v;
        }
        for (int i = 0; i < n; ++i) {
            const Obj* o = &e->o[idx[i]];
            for (int c = 0; c < 4; ++c)
                if (o->contents & (1 << c)) {
                    if (c < 3) st[c] = (o->chopped >> c) & 1;
                    dx[c] = o->x - e->ax[k]; dy[c] = o->y - e->ay[k];
                }
        }
        for (int c = 0; c < 4; ++c) hid[c] = (abs(dx[c]) + abs(dy[c]) <= b->fow) ? 0 : 1;
    }
    const int* off = b->off;
    if (e->comm[0] >= 0) out[off[0] + e->comm[0]] = 1.0;
    out[off[1]] = blind ? 0 : e->ax[0]; out[off[1] + 1] = blind ? 0 : e->ay[0];
    if (e->comm[1] >= 0) out[off[2] + e->comm[1]] = 1.0;
    out[off[3]] = blind ? 0 : e->ax[1]; out[off[3] + 1] = blind ? 0 : e->ay[1];
    int holding = 0;
    for (int s = 0; s < b->nobj; ++s) if (e->o[s].alive && e->o[s].held_by == k) holding = 1;
    out[off[4]] = b->blind[0] ? 0 : holding;                               /* ego BLIND regardless of k (:154) */
    for (int i = 0; i < b->S; ++i) out[off[5] + i] = e->completed[i];
    for (int c = 0; c < 4; ++c) {
        int near = abs(dx[c]) + abs(dy[c]) <= b->fow;
        out[off[6] + c] = hid[c];
        out[off[7] + c] = near ? 0 : dx[c];
        out[off[8] + c] = near ? 0 : dy[c];
        out[off[9] + c] = st[c];
    }
    out[off[10]] = (double)e->t / (double)b->T;
}

/* ---------------------------------------------------------------------------------- API */
Batch* oco_create(const char* level_text, const char* subtasks, int n_envs, int num_agents, int T, int comm_on,
                  int C, int ego_led, int fow, const int* agent_flags /* [A][3] can_move, allergic, blind */,
                  uint64_t seed) {
    Batch* b = (Batch*)calloc(1, sizeof(Batch));
    b->N = n_envs; b->A = num_agents; b->T = T; b->C = C; b->comm_on = comm_on; b->ego_led = ego_led; b->fow = fow; b->seed = seed;
    for (int k = 0; k < num_agents; ++k) { b->can_move[k] = agent_flags[3 * k]; b->allergic[k] = agent_flags[3 * k + 1]; b->blind[k] = agent_flags[3 * k + 2]; }
    if (parse_level(b, level_text) || parse_subtasks(b, subtasks)) { free(b); return NULL; }
    const int sizes[11] = {C, 2, C, 2, 2, b->S, 4, 4, 4, 4, 1};
    int o = 0;
    for (int i = 0; i < 11; ++i) { b->off[i] = o; o += sizes[i]; }
    b->F = o;
    b->env = (Env*)calloc((size_t)n_envs, sizeof(Env));
    for (int i = 0; i < n_envs; ++i) { b->env[i].episodes = 0; b->env[i].comm[0] = 0; b->env[i].comm[1] = 0; env_reset(b, &b->env[i], i, NULL); }
    return b;
}
void oco_destroy(Batch* b) { if (b) { free(b->env); free(b); } }
int oco_obs_width(const Batch* b) { return b->F; }
int oco_num_random(const Batch* b) { return b->nrandom; }
int oco_threads(void) { return host_threads(); }

static void write_obs(const Batch* b, const Env* e, double* obs_env) {
    for (int k = 0; k < b->A; ++k) env_obs(b, e, k, obs_env + (size_t)k * b->F);
}

static void finish_episode(const Batch* b, Env* e, int i) {
    int c = 0;
    for (int s = 0; s < b->S; ++s) c += e->completed[s];
    e->last_completed = c;
    e->episodes += 1;
    env_reset(b, e, i, NULL);
}

typedef struct {
    Batch* b; const uint8_t* mask; const int32_t* placements; const int32_t* actions;
    double *obs, *reward, *term_obs; uint8_t* done; int auto_reset, n_steps; int32_t* actions_out;
    double seconds; long long* totals;
} Job;

static void reset_job(void* ctx, int tid, int nth) {
    Job* j = (Job*)ctx; Batch* b = j->b;
    for (int i = tid; i < b->N; i += nth) {
        if (!j->mask || j->mask[i]) { b->env[i].episodes += 1; env_reset(b, &b->env[i], i, j->placements ? j->placements + (size_t)i * b->nrandom : NULL); }
        if (j->obs) write_obs(b, &b->env[i], j->obs + (size_t)i * b->A * b->F);
    }
}
void oco_reset(Batch* b, const uint8_t* mask, const int32_t* placements, double* obs) {
    Job j = {0}; j.b = b; j.mask = mask; j.placements = placements; j.obs = obs;
    parallel_run(reset_job, &j, b->N);
}

static void step_job(void* ctx, int tid, int nth) {
    Job* j = (Job*)ctx; Batch* b = j->b;
    for (int i = tid; i < b->N; i += nth) {
        Env* e = &b->env[i];
        int nav[MAXA], comm[2] = {0, 0};
        for (int k = 0; k < b->A; ++k) { nav[k] = j->actions[((size_t)i * b->A + k) * 2] & 3; if (k < 2) comm[k] = j->actions[((size_t)i * b->A + k) * 2 + 1]; }
        double r; int d;
        env_step(b, e, nav, comm, &r, &d);
        j->reward[i] = r; j->done[i] = (uint8_t)d;
        if (d && j->auto_reset) {
            if (j->term_obs) write_obs(b, e, j->term_obs + (size_t)i * b->A * b->F);
            finish_episode(b, e, i);
        }
        if (j->obs) write_obs(b, e, j->obs + (size_t)i * b->A * b->F);
    }
}
void oco_step(Batch* b, const int32_t* actions, double* obs, double* reward, uint8_t* done, int auto_reset, double* term_obs) {
    Job j = {0}; j.b = b; j.actions = actions; j.obs = obs; j.reward = reward; j.done = done; j.auto_reset = auto_reset; j.term_obs = term_obs;
    parallel_run(step_job, &j, b->N);
}

/* fused synthetic rollout with the SAME Philox draws as the CUDA oc_rollout */
static void rollout_job(void* ctx, int tid, int nth) {
    Job* j = (Job*)ctx; Batch* b = j->b;
    const size_t AF = (size_t)b->A * b->F;
    for (int i = tid; i < b->N; i += nth) {
        Env* e = &b->env[i];
        for (int s = 0; s < j->n_steps; ++s) {
            uint32_t r[4];
            philox4x32_10((uint32_t)i, b->rollout_step + (uint32_t)s, 0x41435453u, 0u, (uint32_t)b->seed, (uint32_t)(b->seed >> 32), r);
            int nav[MAXA], comm[2];
            for (int k = 0; k < b->A; ++k) nav[k] = (r[0] >> (2 * k)) & 3;
            comm[0] = (int)mulhi(r[1], (uint32_t)b->C); comm[1] = (int)mulhi(r[2], (uint32_t)b->C);
            if (j->actions_out)
                for (int k = 0; k < b->A; ++k) {
                    j->actions_out[(((size_t)s * b->N + i) * b->A + k) * 2] = nav[k];
                    j->actions_out[(((size_t)s * b->N + i) * b->A + k) * 2 + 1] = k < 2 ? comm[k] : 0;
                }
            double rr; int d;
            env_step(b, e, nav, comm, &rr, &d);
            if (j->reward) j->reward[(size_t)s * b->N + i] = rr;
            if (j->done) j->done[(size_t)s * b->N + i] = (uint8_t)d;
            if (d) finish_episode(b, e, i);
            if (j->obs) write_obs(b, e, j->obs + ((size_t)s * b->N + i) * AF);
        }
    }
}
void oco_rollout(Batch* b, int n_steps, double* obs, double* reward, uint8_t* done, int32_t* actions_out) {
    Job j = {0}; j.b = b; j.n_steps = n_steps; j.obs = obs; j.reward = reward; j.done = done; j.actions_out = actions_out;
    parallel_run(rollout_job, &j, b->N);
    b->rollout_step += (uint32_t)n_steps;
}

/* throughput loop for the CPU baseline: random actions from a cheap LCG, obs featurised into a
 * per-thread scratch row (like the reference, which builds fresh obs every step), auto-reset.
 * Returns env-steps done. */
static void throughput_job(void* ctx, int tid, int nth) {
    Job* j = (Job*)ctx; Batch* b = j->b;
    const size_t AF = (size_t)b->A * b->F;
    double* scratch = (double*)malloc(AF * sizeof(double));
    uint64_t lcg = 88172645463325252ull + (uint64_t)tid * 0x9E3779B97F4A7C15ull;
    const double t_end = now_s() + j->seconds;
    long long mine = 0;
    do {
        for (int i = tid; i < b->N; i += nth) {
            Env* e = &b->env[i];
            for (int s = 0; s < 64; ++s) {
                int nav[MAXA], comm[2];
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                uint32_t x = (uint32_t)(lcg >> 32);
                for (int k = 0; k < b->A; ++k) nav[k] = (x >> (2 * k)) & 3;
                comm[0] = (int)mulhi(x * 2654435761u, (uint32_t)b->C); comm[1] = (int)mulhi(x * 40503u + 12345u, (uint32_t)b->C);
                double rr; int d;
                env_step(b, e, nav, comm, &rr, &d);
                if (d) finish_episode(b, e, i);
                write_obs(b, e, scratch);
                mine++;
            }
        }
    } while (now_s() < t_end);
    j->totals[tid] = mine;
    free(scratch);
}
long long oco_throughput(Batch* b, double seconds) {
    long long totals[256] = {0};
    Job j = {0}; j.b = b; j.seconds = seconds; j.totals = totals;
    parallel_run(throughput_job, &j, b->N);
    long long t = 0;
    for (int i = 0; i < 256; ++i) t += totals[i];
    return t;
}

/* canonical state dump for comparisons: per env
 *   ints[0]=t, [1]=episodes, [2..2+A*2) agent xy, then completed bits, count bits (as ints S each),
 *   comm0, comm1, last_completed; objects in world iteration order: nobj rows of
 *   (contents, chopped, x, y, held) padded with -1. */
/* episode clocks from outside (benchmarks stagger them so that a few envs finish in every step) */
void oco_set_clocks(Batch* b, const uint32_t* t) { for (int i = 0; i < b->N; ++i) b->env[i].t = (int)t[i]; }

int oco_state_ints(const Batch* b) { return 2 + 2 * b->A + 2 * b->S + 3 + 5 * b->nobj; }
void oco_get_state(const Batch* b, int32_t* out) {
    const int n = oco_state_ints(b);
    for (int i = 0; i < b->N; ++i) {
        const Env* e = &b->env[i];
        int32_t* o = out + (size_t)i * n;
        int p = 0;
        o[p++] = e->t; o[p++] = e->episodes;
        for (int k = 0; k < b->A; ++k) { o[p++] = e->ax[k]; o[p++] = e->ay[k]; }
        for (int s = 0; s < b->S; ++s) o[p++] = e->completed[s];
        for (int s = 0; s < b->S; ++s) o[p++] = e->count[s];
        o[p++] = e->comm[0]; o[p++] = e->comm[1]; o[p++] = e->last_completed;
        int idx[MAXO], m = 0;
        for (int s = 0; s < b->nobj; ++s) if (e->o[s].alive) idx[m++] = s;
        for (int a = 1; a < m; ++a) {
            int v = idx[a], j = a - 1;
            while (j >= 0) {
                const Obj *x = &e->o[idx[j]], *y = &e->o[v];
                if (e->rank[x->contents] * 256 + x->stamp <= e->rank[y->contents] * 256 + y->stamp) break;
                idx[j + 1] = idx[j]; --j;
            }
            idx[j + 1] = v;
        }
        for (int a = 0; a < b->nobj; ++a) {
            if (a < m) { const Obj* x = &e->o[idx[a]]; o[p++] = x->contents; o[p++] = x->chopped; o[p++] = x->x; o[p++] = x->y; o[p++] = x->held_by >= 0; }
            else { for (int z = 0; z < 5; ++z) o[p++] = -1; }
        }
    }
}
