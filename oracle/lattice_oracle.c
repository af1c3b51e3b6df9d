/*
 * oracle/lattice_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the reference's permutohedral-lattice index path
 * (laoreja/HPLFlowNet, transforms/transforms.py).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs can check (and time) the CUDA path against it.  Nothing under
 * hplflownet_b200/ may import, link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here bit-for-bit against fixtures dumped from the unmodified reference
 * (oracle/make_golden.py, run in the build container where /root/reference
 * exists), and tests/test_oracle_vs_reference.py re-checks live when the
 * reference is importable.
 *
 * Each function cites the reference lines it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HPL_D 3
#define HPL_D1 4

/* transforms/transforms.py:271-276 -- elevate_mat = triu-ones/diag product,
 * evaluated once by torch.mm in fp32.  Bit patterns recorded from that
 * evaluation (row-major 4x3); tests re-derive them. */
static const uint32_t k_elevate_bits[HPL_D1 * HPL_D] = {
    0x3f3504f3u, 0x3ed105ebu, 0x3e93cd3au,
    0xbf3504f3u, 0x3ed105ebu, 0x3e93cd3au,
    0x00000000u, 0xbf5105ebu, 0x3e93cd3au,
    0x00000000u, 0x00000000u, 0xbf5db3d7u};

static inline float bits2f(uint32_t b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}

/* transforms/transforms.py:275 -- expected_std = (d+1)*sqrt(2/3), a Python
 * double; every tensor op that consumes it first narrows it to fp32. */
static const double k_expected_std = 3.265986323710904;

void hplo_constants(float* elevate_4x3, double* expected_std) {
    for (int i = 0; i < HPL_D1 * HPL_D; ++i) elevate_4x3[i] = bits2f(k_elevate_bits[i]);
    *expected_std = k_expected_std;
}

/* transforms/transforms.py:300-353 get_keys_and_barycentric.
 *   pc      (3, n) fp32 row-major
 *   bary    (4, n) fp32      barycentric weights           (:340-345)
 *   emg     (4, n) fp32      elevated - greedy after fix-up (:337)
 *   keys    (4, n, 4) int64  keys[i][p][r] = coord i of the r-th simplex vertex (:347)
 * Arithmetic pinned against the reference: the 4x3 matmul is a k-ordered
 * FMA chain, the scale by expected_std is a separate fp32 multiply, rounding
 * is half-to-even, the descending sort is stable. */
void hplo_keys_barycentric(const float* pc, int64_t n, float* bary, float* emg, int64_t* keys) {
    float E[HPL_D1][HPL_D];
    for (int i = 0; i < HPL_D1; ++i)
        for (int k = 0; k < HPL_D; ++k) E[i][k] = bits2f(k_elevate_bits[i * HPL_D + k]);
    const float std32 = (float)k_expected_std;

    for (int64_t p = 0; p < n; ++p) {
        const float x = pc[p], y = pc[n + p], z = pc[2 * n + p];
        float el[HPL_D1], gr[HPL_D1], em[HPL_D1];
        int rank[HPL_D1];
        for (int i = 0; i < HPL_D1; ++i) {
            float acc = fmaf(E[i][0], x, 0.0f);     /* :309 matmul: FMA chain from a +0 accumulator */
            acc = fmaf(E[i][1], y, acc);
            acc = fmaf(E[i][2], z, acc);
            el[i] = acc * std32;                    /* ... * expected_std         */
            gr[i] = rintf(el[i] / 4.0f) * 4.0f;     /* :312 round-half-even       */
            em[i] = el[i] - gr[i];                  /* :314                       */
        }
        /* :315-319 inverse permutation of a stable descending sort */
        for (int i = 0; i < HPL_D1; ++i) {
            int r = 0;
            for (int j = 0; j < HPL_D1; ++j)
                r += (em[j] > em[i]) || (em[j] == em[i] && j < i);
            rank[i] = r;
        }
        /* :322 */
        const float rsum = (((gr[0] + gr[1]) + gr[2]) + gr[3]) / 4.0f;
        /* :324-334 walk back onto the hyperplane */
        const float sign = rsum > 0.f ? -1.f : (rsum < 0.f ? 1.f : 0.f);
        for (int i = 0; i < HPL_D1; ++i) {
            const float rf = (float)rank[i];
            const int cond = ((rf >= 4.0f - rsum) && rsum > 0.f) || ((rf < -rsum) && rsum < 0.f);
            const float step = 4.0f * sign * (float)cond;
            gr[i] += step;
            rank[i] += (int)step;
            rank[i] += (int)rsum;
        }
        /* :337-345 */
        float b[HPL_D1 + 1] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < HPL_D1; ++i) em[i] = el[i] - gr[i];
        for (int i = 0; i < HPL_D1; ++i) b[HPL_D - rank[i]] += em[i];
        for (int i = 0; i < HPL_D1; ++i) b[HPL_D1 - rank[i]] -= em[i];
        for (int j = 0; j <= HPL_D1; ++j) b[j] /= 4.0f;
        b[0] += 1.0f + b[HPL_D1];
        for (int j = 0; j < HPL_D1; ++j) bary[j * n + p] = b[j];
        for (int i = 0; i < HPL_D1; ++i) emg[i * n + p] = em[i];
        /* :281-285,:347 canonical[rank][r] = r if rank < d1-r else r-d1 */
        for (int i = 0; i < HPL_D1; ++i) {
            const int64_t g = (int64_t)gr[i];
            for (int r = 0; r < HPL_D1; ++r)
                keys[(i * n + p) * HPL_D1 + r] = g + (rank[i] < HPL_D1 - r ? r : r - HPL_D1);
        }
    }
}

/* transforms/transforms.py:384-385 -- per-coordinate key range over a cloud.
 * mins/maxs are updated in place so the caller can fold two clouds. */
void hplo_key_range(const int64_t* keys, int64_t n, int64_t* mins, int64_t* maxs) {
    for (int i = 0; i < HPL_D1; ++i)
        for (int64_t q = 0; q < n * HPL_D1; ++q) {
            const int64_t v = keys[i * n * HPL_D1 + q];
            if (v < mins[i]) mins[i] = v;
            if (v > maxs[i]) maxs[i] = v;
        }
}

/* transforms/transforms.py:70-86 key2int (mixed radix, radix 0 unused, no range check) */
static inline int64_t pack_key(const int64_t* key, const int64_t* mins, const int64_t* maxs) {
    int64_t res = 0;
    for (int i = 0; i < HPL_D; ++i) {
        res += key[i] - mins[i];
        res *= maxs[i + 1] - mins[i + 1] + 1;
    }
    return res + (key[HPL_D] - mins[HPL_D]);
}

/* transforms/transforms.py:89-100 int2key (Python % and // : floor semantics) */
static inline void unpack_key(int64_t v, const int64_t* mins, const int64_t* maxs, int64_t* key) {
    for (int i = HPL_D; i > 0; --i) {
        const int64_t s = maxs[i] - mins[i] + 1;
        int64_t m = v % s;
        if (m != 0 && ((m < 0) != (s < 0))) m += s;
        key[i] = m;
        v = (v - m) / s; /* exact: v-m is a multiple of s */
    }
    key[0] = v;
    for (int i = 0; i < HPL_D1; ++i) key[i] += mins[i];
}

/* A minimal int64 -> int64 map (stand-in for the reference's khash table,
 * models/khash_int2int.h:8-33).  Only the map semantics matter for parity --
 * vertex ids are the insertion order, not bucket positions. */
typedef struct {
    int64_t* k;
    int64_t* v;
    uint8_t* used;
    uint64_t cap, cnt;
} omap;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    return x ^ (x >> 33);
}
static void omap_init(omap* m, uint64_t want) {
    m->cap = 64;
    while (m->cap < want * 2) m->cap <<= 1;
    m->cnt = 0;
    m->k = (int64_t*)malloc(m->cap * 8);
    m->v = (int64_t*)malloc(m->cap * 8);
    m->used = (uint8_t*)calloc(m->cap, 1);
}
static void omap_free(omap* m) { free(m->k); free(m->v); free(m->used); }
static int64_t omap_get(const omap* m, int64_t key, int64_t dflt) {
    uint64_t i = mix64((uint64_t)key) & (m->cap - 1);
    while (m->used[i]) {
        if (m->k[i] == key) return m->v[i];
        i = (i + 1) & (m->cap - 1);
    }
    return dflt;
}
static void omap_put(omap* m, int64_t key, int64_t val) {
    uint64_t i = mix64((uint64_t)key) & (m->cap - 1);
    while (m->used[i]) {
        if (m->k[i] == key) { m->v[i] = val; return; }
        i = (i + 1) & (m->cap - 1);
    }
    m->used[i] = 1; m->k[i] = key; m->v[i] = val; m->cnt++;
}

/* transforms/transforms.py:179-207 -- first-occurrence vertex ids, scanning
 * point outer / remainder inner.  vert_keys (cap 4n) receives the packed key
 * of each new vertex (the reference's key_hash_table, :186); last_pc (4, hcap)
 * receives the vertex coordinates when non-NULL (:188-189).  Returns H. */
static int64_t insert_cloud(omap* m, const int64_t* keys, int64_t n, const int64_t* mins,
                            const int64_t* maxs, int64_t* lattice_offset, int64_t* vert_keys,
                            float* last_pc, int64_t last_ld) {
    int64_t h = 0;
    for (int64_t p = 0; p < n; ++p)
        for (int r = 0; r < HPL_D1; ++r) {
            int64_t key[HPL_D1];
            for (int i = 0; i < HPL_D1; ++i) key[i] = keys[(i * n + p) * HPL_D1 + r];
            const int64_t packed = pack_key(key, mins, maxs);
            int64_t id = omap_get(m, packed, -1);
            if (id == -1) {
                id = h++;
                omap_put(m, packed, id);
                vert_keys[id] = packed;
                if (last_pc)
                    for (int i = 0; i < HPL_D1; ++i) last_pc[i * last_ld + id] = (float)key[i];
            }
            lattice_offset[r * n + p] = id;
        }
    return h;
}

/* transforms/transforms.py:387-391 -- H = number of distinct simplex vertices.
 * (The reference counts with a Python set before allocating; same number.) */
int64_t hplo_count_vertices(const int64_t* keys, int64_t n, const int64_t* mins, const int64_t* maxs) {
    omap m;
    omap_init(&m, (uint64_t)n * HPL_D1);
    for (int64_t p = 0; p < n; ++p)
        for (int r = 0; r < HPL_D1; ++r) {
            int64_t key[HPL_D1];
            for (int i = 0; i < HPL_D1; ++i) key[i] = keys[(i * n + p) * HPL_D1 + r];
            omap_put(&m, pack_key(key, mins, maxs), 1);
        }
    const int64_t h = (int64_t)m.cnt;
    omap_free(&m);
    return h;
}

/* transforms/transforms.py:133-261 build_unsymmetric.
 * Sizes: bcn_fs / corr_fs / corr_cs = filter sizes or -1 ("do not build").
 * Offsets tables are (size, 4) int64.  Output tables must be pre-filled by the
 * caller exactly like the reference does (-1, :398-415).
 *   blur1 (bcn_fs, h1)  blur2 (bcn_fs, h2)
 *   corr1 (corr_cs, h1) corr2 (corr_fs, corr_cs, h1)
 *   last1 (4, h1) last2 (4, h2) fp32 or NULL (assign_last false, :188,:203)
 * h1/h2 are the expected vertex counts (leading dimensions). Returns 0, or -1
 * if the counts found differ from h1/h2. */
int hplo_build_unsymmetric(int64_t n1, int64_t n2, int64_t bcn_fs, int64_t corr_fs, int64_t corr_cs,
                           const int64_t* keys1, const int64_t* keys2, const int64_t* maxs,
                           const int64_t* mins, int64_t* off1, int64_t* off2,
                           const int64_t* bcn_offsets, int64_t* blur1, int64_t* blur2,
                           const int64_t* corr_f_offsets, const int64_t* corr_c_offsets,
                           int64_t* corr1, int64_t* corr2, float* last1, float* last2, int64_t h1,
                           int64_t h2) {
    omap t1, t2;
    omap_init(&t1, (uint64_t)n1 * HPL_D1);
    omap_init(&t2, (uint64_t)n2 * HPL_D1);
    int64_t* vk1 = (int64_t*)malloc((size_t)(n1 * HPL_D1 + 1) * 8);
    int64_t* vk2 = (int64_t*)malloc((size_t)(n2 * HPL_D1 + 1) * 8);
    const int64_t c1 = insert_cloud(&t1, keys1, n1, mins, maxs, off1, vk1, last1, h1);
    const int64_t c2 = insert_cloud(&t2, keys2, n2, mins, maxs, off2, vk2, last2, h2);
    int rc = (c1 == h1 && c2 == h2) ? 0 : -1;

    if (rc == 0) {
        for (int64_t h = 0; h < c1; ++h) { /* :209-241 */
            int64_t key[HPL_D1], nk[HPL_D1], nk2[HPL_D1];
            unpack_key(vk1[h], mins, maxs, key);
            if (bcn_fs != -1)
                for (int64_t f = 0; f < bcn_fs; ++f) {
                    for (int i = 0; i < HPL_D1; ++i) nk[i] = key[i] + bcn_offsets[f * HPL_D1 + i];
                    blur1[f * h1 + h] = omap_get(&t1, pack_key(nk, mins, maxs), -1);
                }
            if (corr_fs != -1)
                for (int64_t c = 0; c < corr_cs; ++c) {
                    for (int i = 0; i < HPL_D1; ++i) nk[i] = key[i] + corr_c_offsets[c * HPL_D1 + i];
                    corr1[c * h1 + h] = omap_get(&t1, pack_key(nk, mins, maxs), -1);
                    for (int64_t f = 0; f < corr_fs; ++f) {
                        for (int i = 0; i < HPL_D1; ++i) nk2[i] = nk[i] + corr_f_offsets[f * HPL_D1 + i];
                        corr2[(f * corr_cs + c) * h1 + h] = omap_get(&t2, pack_key(nk2, mins, maxs), -1);
                    }
                }
        }
        if (bcn_fs != -1) /* :243-255 */
            for (int64_t h = 0; h < c2; ++h) {
                int64_t key[HPL_D1], nk[HPL_D1];
                unpack_key(vk2[h], mins, maxs, key);
                for (int64_t f = 0; f < bcn_fs; ++f) {
                    for (int i = 0; i < HPL_D1; ++i) nk[i] = key[i] + bcn_offsets[f * HPL_D1 + i];
                    blur2[f * h2 + h] = omap_get(&t2, pack_key(nk, mins, maxs), -1);
                }
            }
    }
    free(vk1); free(vk2);
    omap_free(&t1); omap_free(&t2);
    return rc;
}

/* transforms/transforms.py:461-467 -- positions of this scale's vertices, the
 * next scale's input points:  p = E^T . (key / fp32(expected_std*scale)).
 * True fp32 division; the 3x4 matmul is a k-ordered FMA chain.
 *   last (4, h) fp32 vertex coordinates   ->   out (3, h) fp32 */
void hplo_next_points(const float* last, int64_t h, double scale, float* out) {
    float E[HPL_D1][HPL_D];
    for (int i = 0; i < HPL_D1; ++i)
        for (int k = 0; k < HPL_D; ++k) E[i][k] = bits2f(k_elevate_bits[i * HPL_D + k]);
    const float div = (float)(k_expected_std * scale);
    for (int64_t q = 0; q < h; ++q) {
        float v[HPL_D1];
        for (int i = 0; i < HPL_D1; ++i) v[i] = last[i * h + q] / div;
        for (int k = 0; k < HPL_D; ++k) {
            float acc = 0.0f;
            for (int i = 0; i < HPL_D1; ++i) acc = fmaf(E[i][k], v[i], acc);
            out[k * h + q] = acc;
        }
    }
}

/* transforms/transforms.py:103-130,:292-298 -- Traverse.go enumerates the
 * neighbourhood offsets of a given radius in the order the conv weights index
 * them.  out must hold ((r+1)^4 - r^4) * 4 int64.  Returns the count. */
static void walk(int radius, const int64_t* start, int d, int has_zero, int64_t* out, int64_t* cnt) {
    if (d > HPL_D) {
        memcpy(out + (*cnt) * HPL_D1, start, sizeof(int64_t) * HPL_D1);
        (*cnt)++;
        return;
    }
    int64_t cur[HPL_D1];
    memcpy(cur, start, sizeof(cur));
    const int range_end = (has_zero || d < HPL_D) ? radius + 1 : 1;
    for (int i = 0; i < range_end; ++i) {
        walk(radius, cur, d + 1, has_zero || i == 0, out, cnt);
        /* advance_in_dimension(d1, 1, d, key): key -= 1; key[d] += d1 */
        for (int j = 0; j < HPL_D1; ++j) cur[j] -= 1;
        cur[d] += HPL_D1;
    }
}
int64_t hplo_neighbor_offsets(int radius, int64_t* out) {
    const int64_t origin[HPL_D1] = {0, 0, 0, 0};
    int64_t cnt = 0;
    walk(radius, origin, 0, 0, out, &cnt);
    return cnt;
}
