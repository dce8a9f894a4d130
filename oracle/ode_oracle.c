/* ode_oracle.c -- TEST INFRASTRUCTURE ONLY (see ode_oracle.h: "parity unpinned").
 *
 * Plain C99 float32 restatement of the libode slice behind the reference's physics tick
 * (/root/reference/src/main.c:206-216 -> dSpaceCollide / NearCallback :674-693 / dWorldStep).
 * libode itself is an un-vendored dependency (src/main.c:11); every routine below restates the
 * published ODE 0.13-0.16 (dSINGLE) algorithm named in its comment, see SURVEY.md Appendix A.
 *
 * Compile with -ffp-contract=off so no FMA contraction happens: the CUDA path is compiled with
 * --fmad=false and uses the same operation order in its predicates.
 */
#include "ode_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INF INFINITY

/* ------------------------------------------------------------------ small math (ODE odemath.h) */

static float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* dCalcVectorDot3_14: a stride 1, b stride 4 (a column of a 3x4 matrix) */
static float dot3_14(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[4] + a[2] * b[8]; }
static float dot3_41(const float *a, const float *b) { return a[0] * b[0] + a[4] * b[1] + a[8] * b[2]; }
static float dot3_44(const float *a, const float *b) { return a[0] * b[0] + a[4] * b[4] + a[8] * b[8]; }
/* dCalcVectorCross3: a = b x c */
static void cross3(float *a, const float *b, const float *c) {
    a[0] = b[1] * c[2] - b[2] * c[1];
    a[1] = b[2] * c[0] - b[0] * c[2];
    a[2] = b[0] * c[1] - b[1] * c[0];
}
/* dMultiply0_331: a = B(3x4) * c */
static void mul0_331(float *a, const float *B, const float *c) {
    a[0] = dot3(B, c);
    a[1] = dot3(B + 4, c);
    a[2] = dot3(B + 8, c);
}
/* dMultiply1_331: a = B^T * c */
static void mul1_331(float *a, const float *B, const float *c) {
    a[0] = dot3_41(B, c);
    a[1] = dot3_41(B + 1, c);
    a[2] = dot3_41(B + 2, c);
}
/* dMultiply0_333: A = B*C ; dMultiply2_333: A = B*C^T (all 3x4 row-major) */
static void mul0_333(float *A, const float *B, const float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i * 4 + j] = dot3_14(B + i * 4, C + j);
}
static void mul2_333(float *A, const float *B, const float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i * 4 + j] = dot3(B + i * 4, C + j * 4);
}

/* dSafeNormalize3 (odemath.cpp): scale by the largest component first */
static void normalize3(float *a) {
    float aa0 = fabsf(a[0]), aa1 = fabsf(a[1]), aa2 = fabsf(a[2]);
    int idx;
    if (aa1 > aa0) {
        idx = (aa2 > aa1) ? 2 : 1;
    } else if (aa2 > aa0) {
        idx = 2;
    } else {
        if (aa0 <= 0) { a[0] = 1; a[1] = 0; a[2] = 0; return; }
        idx = 0;
    }
    float s = (idx == 0) ? aa0 : (idx == 1 ? aa1 : aa2);
    a[0] /= s; a[1] /= s; a[2] /= s;
    float l = 1.0f / sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    a[0] *= l; a[1] *= l; a[2] *= l;
}

/* dSafeNormalize4 */
static void normalize4(float *a) {
    float l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0) {
        l = 1.0f / sqrtf(l);
        a[0] *= l; a[1] *= l; a[2] *= l; a[3] *= l;
    } else {
        a[0] = 1; a[1] = 0; a[2] = 0; a[3] = 0;
    }
}

/* dQtoR (rotation.cpp) */
void orc_q_to_r(const float q[4], float R[12]) {
    float qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3;
    R[1] = 2 * (q[1] * q[2] - q[0] * q[3]);
    R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[3] = 0;
    R[4] = 2 * (q[1] * q[2] + q[0] * q[3]);
    R[5] = 1 - qq1 - qq3;
    R[6] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[7] = 0;
    R[8] = 2 * (q[1] * q[3] - q[0] * q[2]);
    R[9] = 2 * (q[2] * q[3] + q[0] * q[1]);
    R[10] = 1 - qq1 - qq2;
    R[11] = 0;
}

/* dRtoQ (rotation.cpp) */
void orc_r_to_q(const float R[12], float q[4]) {
    float tr = R[0] + R[5] + R[10], s;
    if (tr >= 0) {
        s = sqrtf(tr + 1);
        q[0] = 0.5f * s;
        s = 0.5f / s;
        q[1] = (R[9] - R[6]) * s;
        q[2] = (R[2] - R[8]) * s;
        q[3] = (R[4] - R[1]) * s;
    } else {
        int c;
        if (R[5] > R[0]) c = (R[10] > R[5]) ? 2 : 1;
        else c = (R[10] > R[0]) ? 2 : 0;
        if (c == 0) {
            s = sqrtf((R[0] - (R[5] + R[10])) + 1);
            q[1] = 0.5f * s;
            s = 0.5f / s;
            q[2] = (R[1] + R[4]) * s;
            q[3] = (R[8] + R[2]) * s;
            q[0] = (R[9] - R[6]) * s;
        } else if (c == 1) {
            s = sqrtf((R[5] - (R[10] + R[0])) + 1);
            q[2] = 0.5f * s;
            s = 0.5f / s;
            q[3] = (R[6] + R[9]) * s;
            q[1] = (R[1] + R[4]) * s;
            q[0] = (R[2] - R[8]) * s;
        } else {
            s = sqrtf((R[10] - (R[0] + R[5])) + 1);
            q[3] = 0.5f * s;
            s = 0.5f / s;
            q[1] = (R[8] + R[2]) * s;
            q[2] = (R[6] + R[9]) * s;
            q[0] = (R[4] - R[1]) * s;
        }
    }
}

/* dPlaneSpace (odemath.cpp) */
void orc_plane_space(const float n[3], float p[3], float q[3]) {
    if (fabsf(n[2]) > 0.70710678118654752440f) {
        float a = n[1] * n[1] + n[2] * n[2];
        float k = 1.0f / sqrtf(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        float a = n[0] * n[0] + n[1] * n[1];
        float k = 1.0f / sqrtf(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

/* reference PRNG, src/rand.c:7-13 (Weyl + two multiply-xorshift rounds) */
unsigned orc_rand_next(unsigned *state) {
    *state += 0xE120FC15u;
    unsigned long long t = (unsigned long long)(*state) * 0x4A39B70Dull;
    unsigned m1 = (unsigned)((t >> 32) ^ t);
    t = (unsigned long long)m1 * 0x12FAD5C9ull;
    return (unsigned)((t >> 32) ^ t);
}

/* ------------------------------------------------------------------ world state */

typedef struct {
    float pos[3], q[4], R[12], lvel[3], avel[3];
    float mass, I[12], invMass, invI[12];
    float facc[3], tacc[3];
    int flags, env;
} obody;

typedef struct {
    int type, body, env, mesh;
    float dims[4];
    float pos[3], R[12]; /* static pose (or cache of the body pose) */
    unsigned cat, col;
    float aabb[6]; /* minx maxx miny maxy minz maxz (ODE order) */
} ogeom;

typedef struct {
    float *v; /* 3*nv */
    int *t;   /* 3*nt */
    int nv, nt;
    float lo[3], hi[3];
} omesh;

typedef struct {
    orc_contact_geom g;
    orc_surface s;
    int b1, b2; /* node[0], node[1] after dJointAttach's swap; b1 may be -1 only if both are */
    int reverse;
} ojoint;

struct orc_world {
    float gravity[3], erp, cfm, sor_w, max_vel, min_depth;
    int iters;
    obody *b; int nb, capb;
    ogeom *g; int ng, capg;
    omesh *m; int nm, capm;
    ojoint *j; int nj, capj;
    unsigned long lcg_seed; /* ODE dRand state */
    int perturb_fma;        /* experiment: FMA-contracted solver arithmetic (not ODE's rounding) */
    float *last_lambda; int nrows;
    /* diagnostics kept for the tests (orc_set_keep_rows): the last step's rows before Ad scaling */
    int keep_rows;
    float *dbg_J, *dbg_c, *dbg_cfm, *dbg_lo, *dbg_hi, *dbg_rhs; int *dbg_jb; int dbg_m;
};

orc_world *orc_create(void) {
    orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
    /* dWorldCreate defaults (ode.cpp): ERP 0.2, CFM 1e-5 (single), QuickStep 20 iterations, w 1.3 */
    w->erp = 0.2f; w->cfm = 1e-5f; w->sor_w = 1.3f; w->iters = 20;
    w->max_vel = ORC_INF; w->min_depth = 0;
    return w;
}
static void dbg_free(orc_world *w) {
    free(w->dbg_J); free(w->dbg_c); free(w->dbg_cfm); free(w->dbg_lo); free(w->dbg_hi); free(w->dbg_rhs); free(w->dbg_jb);
    w->dbg_J = w->dbg_c = w->dbg_cfm = w->dbg_lo = w->dbg_hi = w->dbg_rhs = 0; w->dbg_jb = 0; w->dbg_m = 0;
}
void orc_destroy(orc_world *w) {
    if (!w) return;
    dbg_free(w);
    for (int i = 0; i < w->nm; i++) { free(w->m[i].v); free(w->m[i].t); }
    free(w->b); free(w->g); free(w->m); free(w->j); free(w->last_lambda); free(w);
}
void orc_set_gravity(orc_world *w, float x, float y, float z) { w->gravity[0] = x; w->gravity[1] = y; w->gravity[2] = z; }
void orc_set_params(orc_world *w, float erp, float cfm, int iters, float sor_w) {
    w->erp = erp; w->cfm = cfm; w->iters = iters; w->sor_w = sor_w;
}
void orc_set_perturb_fma(orc_world *w, int on) { w->perturb_fma = on; }
void orc_set_contact_params(orc_world *w, float max_vel, float min_depth) { w->max_vel = max_vel; w->min_depth = min_depth; }

static void invert3_sym(const float *I, float *inv) {
    /* general 3x3 inverse via cofactors (dInvertPDMatrix's result for the PD inertia tensor) */
    float a = I[0], b = I[1], c = I[2], d = I[4], e = I[5], f = I[6], g = I[8], h = I[9], i = I[10];
    float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    float det = a * A + b * B + c * C;
    float id = 1.0f / det;
    memset(inv, 0, 12 * sizeof(float));
    inv[0] = A * id; inv[1] = -(b * i - c * h) * id; inv[2] = (b * f - c * e) * id;
    inv[4] = B * id; inv[5] = (a * i - c * g) * id; inv[6] = -(a * f - c * d) * id;
    inv[8] = C * id; inv[9] = -(a * h - b * g) * id; inv[10] = (a * e - b * d) * id;
}

int orc_add_body(orc_world *w, const float pos[3], const float q[4], const float *R12,
                 const float lvel[3], const float avel[3], float mass, const float *inertia9,
                 int flags, int env) {
    if (w->nb == w->capb) {
        w->capb = w->capb ? w->capb * 2 : 64;
        w->b = (obody *)realloc(w->b, sizeof(obody) * (size_t)w->capb);
    }
    obody *b = &w->b[w->nb];
    memset(b, 0, sizeof(*b));
    memcpy(b->pos, pos, 12);
    if (q) memcpy(b->q, q, 16); else { b->q[0] = 1; }
    if (R12) {
        /* dBodySetRotation: copy R, q = dRtoQ(R), normalise */
        memcpy(b->R, R12, 48); b->R[3] = b->R[7] = b->R[11] = 0;
        orc_r_to_q(b->R, b->q);
        normalize4(b->q);
    } else {
        /* dBodySetQuaternion: normalise, R = dQtoR(q) */
        normalize4(b->q);
        orc_q_to_r(b->q, b->R);
    }
    if (lvel) memcpy(b->lvel, lvel, 12);
    if (avel) memcpy(b->avel, avel, 12);
    b->mass = mass;
    if (inertia9) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) b->I[i * 4 + j] = inertia9[i * 3 + j];
    } else {
        b->I[0] = b->I[5] = b->I[10] = 1;
    }
    b->flags = flags; b->env = env;
    if (flags & ORC_BODY_KINEMATIC) {
        /* dBodySetKinematic: invMass = 0, invI = 0 */
        b->invMass = 0; memset(b->invI, 0, sizeof(b->invI));
    } else {
        b->invMass = 1.0f / mass;
        invert3_sym(b->I, b->invI);
    }
    return w->nb++;
}
int orc_num_bodies(const orc_world *w) { return w->nb; }
void orc_get_body(const orc_world *w, int i, float pos[3], float q[4], float R12[12], float lvel[3], float avel[3]) {
    const obody *b = &w->b[i];
    if (pos) memcpy(pos, b->pos, 12);
    if (q) memcpy(q, b->q, 16);
    if (R12) memcpy(R12, b->R, 48);
    if (lvel) memcpy(lvel, b->lvel, 12);
    if (avel) memcpy(avel, b->avel, 12);
}
void orc_set_body_state(orc_world *w, int i, const float pos[3], const float q[4], const float *R12,
                        const float lvel[3], const float avel[3]) {
    obody *b = &w->b[i];
    if (pos) memcpy(b->pos, pos, 12);
    if (q) { memcpy(b->q, q, 16); if (!R12) orc_q_to_r(b->q, b->R); }
    if (R12) { memcpy(b->R, R12, 48); b->R[3] = b->R[7] = b->R[11] = 0; }
    if (lvel) memcpy(b->lvel, lvel, 12);
    if (avel) memcpy(b->avel, avel, 12);
}
void orc_add_force(orc_world *w, int i, const float f[3], const float t[3]) {
    obody *b = &w->b[i];
    if (f) for (int k = 0; k < 3; k++) b->facc[k] += f[k];
    if (t) for (int k = 0; k < 3; k++) b->tacc[k] += t[k];
}

int orc_add_mesh(orc_world *w, const float *verts, int nverts, const int *tris, int ntris) {
    if (w->nm == w->capm) {
        w->capm = w->capm ? w->capm * 2 : 4;
        w->m = (omesh *)realloc(w->m, sizeof(omesh) * (size_t)w->capm);
    }
    omesh *m = &w->m[w->nm];
    m->nv = nverts; m->nt = ntris;
    m->v = (float *)malloc(sizeof(float) * 3 * (size_t)nverts);
    m->t = (int *)malloc(sizeof(int) * 3 * (size_t)ntris);
    memcpy(m->v, verts, sizeof(float) * 3 * (size_t)nverts);
    memcpy(m->t, tris, sizeof(int) * 3 * (size_t)ntris);
    for (int k = 0; k < 3; k++) { m->lo[k] = ORC_INF; m->hi[k] = -ORC_INF; }
    for (int i = 0; i < nverts; i++)
        for (int k = 0; k < 3; k++) {
            m->lo[k] = fminf(m->lo[k], verts[3 * i + k]);
            m->hi[k] = fmaxf(m->hi[k], verts[3 * i + k]);
        }
    return w->nm++;
}

int orc_add_geom(orc_world *w, int type, const float dims[4], int body, const float pos[3],
                 const float *R12, unsigned cat, unsigned col, int env) {
    if (w->ng == w->capg) {
        w->capg = w->capg ? w->capg * 2 : 64;
        w->g = (ogeom *)realloc(w->g, sizeof(ogeom) * (size_t)w->capg);
    }
    ogeom *g = &w->g[w->ng];
    memset(g, 0, sizeof(*g));
    g->type = type; g->body = body; g->env = env; g->cat = cat; g->col = col;
    memcpy(g->dims, dims, 16);
    g->R[0] = g->R[5] = g->R[10] = 1;
    if (pos) memcpy(g->pos, pos, 12);
    if (R12) { memcpy(g->R, R12, 48); g->R[3] = g->R[7] = g->R[11] = 0; }
    if (type == ORC_PLANE) {
        /* dCreatePlane -> make_sure_plane_normal_has_unit_length */
        float l = dims[0] * dims[0] + dims[1] * dims[1] + dims[2] * dims[2];
        if (l > 0) {
            l = 1.0f / sqrtf(l);
            for (int k = 0; k < 4; k++) g->dims[k] = dims[k] * l;
        } else {
            g->dims[0] = 1; g->dims[1] = 0; g->dims[2] = 0; g->dims[3] = 0;
        }
    }
    if (type == ORC_TRIMESH) g->mesh = (int)dims[0];
    return w->ng++;
}
int orc_num_geoms(const orc_world *w) { return w->ng; }

/* ------------------------------------------------------------------ AABBs (ODE computeAABB) */

static void geom_sync_pose(orc_world *w, ogeom *g) {
    if (g->body >= 0) {
        memcpy(g->pos, w->b[g->body].pos, 12);
        memcpy(g->R, w->b[g->body].R, 48);
    }
}

static void geom_aabb(orc_world *w, ogeom *g) {
    geom_sync_pose(w, g);
    float *a = g->aabb;
    if (g->type == ORC_SPHERE) {
        /* dxSphere::computeAABB */
        float r = g->dims[0];
        a[0] = g->pos[0] - r; a[1] = g->pos[0] + r;
        a[2] = g->pos[1] - r; a[3] = g->pos[1] + r;
        a[4] = g->pos[2] - r; a[5] = g->pos[2] + r;
    } else if (g->type == ORC_BOX) {
        /* dxBox::computeAABB: range_i = 0.5 * sum_j |R[i][j] * side_j| */
        const float *R = g->R;
        for (int i = 0; i < 3; i++) {
            float range = 0.5f * (fabsf(R[i * 4 + 0] * g->dims[0]) + fabsf(R[i * 4 + 1] * g->dims[1]) +
                                  fabsf(R[i * 4 + 2] * g->dims[2]));
            a[2 * i] = g->pos[i] - range;
            a[2 * i + 1] = g->pos[i] + range;
        }
    } else if (g->type == ORC_PLANE) {
        /* dxPlane::computeAABB: infinite, half-space if axis aligned */
        const float *p = g->dims;
        a[0] = -ORC_INF; a[1] = ORC_INF; a[2] = -ORC_INF; a[3] = ORC_INF; a[4] = -ORC_INF; a[5] = ORC_INF;
        if (p[1] == 0.0f && p[2] == 0.0f) {
            a[0] = (p[0] > 0) ? -ORC_INF : -p[3];
            a[1] = (p[0] > 0) ? p[3] : ORC_INF;
        } else if (p[0] == 0.0f && p[2] == 0.0f) {
            a[2] = (p[1] > 0) ? -ORC_INF : -p[3];
            a[3] = (p[1] > 0) ? p[3] : ORC_INF;
        } else if (p[0] == 0.0f && p[1] == 0.0f) {
            a[4] = (p[2] > 0) ? -ORC_INF : -p[3];
            a[5] = (p[2] > 0) ? p[3] : ORC_INF;
        }
    } else if (g->type == ORC_TRIMESH) {
        /* transformed mesh-local box: centre + |R| * half-extent (engine's own rule) */
        const omesh *m = &w->m[g->mesh];
        float c[3], e[3];
        for (int k = 0; k < 3; k++) { c[k] = 0.5f * (m->lo[k] + m->hi[k]); e[k] = 0.5f * (m->hi[k] - m->lo[k]); }
        const float *R = g->R;
        for (int i = 0; i < 3; i++) {
            float wc = g->pos[i] + (R[i * 4 + 0] * c[0] + R[i * 4 + 1] * c[1] + R[i * 4 + 2] * c[2]);
            float range = fabsf(R[i * 4 + 0] * e[0]) + fabsf(R[i * 4 + 1] * e[1]) + fabsf(R[i * 4 + 2] * e[2]);
            a[2 * i] = wc - range;
            a[2 * i + 1] = wc + range;
        }
    }
}

void orc_get_aabb(orc_world *w, int g, float aabb[6]) {
    geom_aabb(w, &w->g[g]);
    memcpy(aabb, w->g[g].aabb, 24);
}

/* collideAABBs (collision_space_internal.h) + the env rule of batched worlds */
static int pair_passes(const ogeom *g1, const ogeom *g2) {
    if (g1->body == g2->body && g1->body >= 0) return 0;
    if (g1->env >= 0 && g2->env >= 0 && g1->env != g2->env) return 0;
    if (!((g1->cat & g2->col) || (g2->cat & g1->col))) return 0;
    const float *a = g1->aabb, *b = g2->aabb;
    if (a[0] > b[1] || a[1] < b[0] || a[2] > b[3] || a[3] < b[2] || a[4] > b[5] || a[5] < b[4]) return 0;
    return 1;
}

typedef struct { int *p; long n, cap, written_cap; } pairbuf;
static void pb_push(pairbuf *pb, int a, int b) {
    if (pb->n == pb->cap) {
        pb->cap = pb->cap ? pb->cap * 2 : 1024;
        pb->p = (int *)realloc(pb->p, sizeof(int) * 2 * (size_t)pb->cap);
    }
    pb->p[2 * pb->n] = a < b ? a : b;
    pb->p[2 * pb->n + 1] = a < b ? b : a;
    pb->n++;
}
static int cmp_pair(const void *x, const void *y) {
    const int *a = (const int *)x, *b = (const int *)y;
    if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
    if (a[1] != b[1]) return a[1] < b[1] ? -1 : 1;
    return 0;
}

/* dxHashSpace::collide restated: level = smallest L with 2^L >= largest AABB side, clamped to
 * [-3, 10]; a geom is entered in every cell its AABB touches at its level; it is looked up at
 * its own and every coarser level; infinite AABBs ("big boxes") are tested against everything.
 * ODE dedupes with an n*n bit matrix; here a pair is emitted only from the lowest shared cell. */
typedef struct hnode { int geom, level, c[3]; struct hnode *next; } hnode;

static unsigned long hkey(int level, int x, int y, int z) {
    return ((unsigned long)(unsigned)level * 1000003ul) ^ ((unsigned long)(unsigned)x * 73856093ul) ^
           ((unsigned long)(unsigned)y * 19349663ul) ^ ((unsigned long)(unsigned)z * 83492791ul);
}

static void hash_broadphase(orc_world *w, pairbuf *pb) {
    int n = w->ng;
    int *level = (int *)malloc(sizeof(int) * (size_t)n);
    int(*db)[6] = (int(*)[6])malloc(sizeof(int[6]) * (size_t)n);
    int *big = (int *)malloc(sizeof(int) * (size_t)n);
    int nbig = 0;
    const int minlevel = -3, maxlevel = 10;
    size_t ncells_total = 0;
    for (int i = 0; i < n; i++) {
        const float *a = w->g[i].aabb;
        int inf = 0;
        for (int k = 0; k < 6; k++) if (isinf(a[k])) inf = 1;
        if (inf) { level[i] = 1000; big[nbig++] = i; continue; }
        float maxsize = fmaxf(a[1] - a[0], fmaxf(a[3] - a[2], a[5] - a[4]));
        int L = minlevel;
        while (L < maxlevel && ldexpf(1.0f, L) < maxsize) L++;
        level[i] = L;
        float cs = ldexpf(1.0f, L);
        size_t cells = 1;
        for (int k = 0; k < 3; k++) {
            db[i][2 * k] = (int)floorf(a[2 * k] / cs);
            db[i][2 * k + 1] = (int)floorf(a[2 * k + 1] / cs);
            cells *= (size_t)(db[i][2 * k + 1] - db[i][2 * k] + 1);
        }
        ncells_total += cells;
    }
    size_t hsize = 16;
    while (hsize < ncells_total * 2) hsize <<= 1;
    hnode **table = (hnode **)calloc(hsize, sizeof(hnode *));
    hnode *pool = (hnode *)malloc(sizeof(hnode) * (ncells_total + 1));
    size_t np = 0;
    for (int i = 0; i < n; i++) {
        if (level[i] == 1000) continue;
        for (int x = db[i][0]; x <= db[i][1]; x++)
            for (int y = db[i][2]; y <= db[i][3]; y++)
                for (int z = db[i][4]; z <= db[i][5]; z++) {
                    hnode *nd = &pool[np++];
                    nd->geom = i; nd->level = level[i]; nd->c[0] = x; nd->c[1] = y; nd->c[2] = z;
                    size_t h = hkey(level[i], x, y, z) & (hsize - 1);
                    nd->next = table[h]; table[h] = nd;
                }
    }
    for (int i = 0; i < n; i++) {
        if (level[i] == 1000) continue;
        int b[6];
        memcpy(b, db[i], sizeof(b));
        for (int L = level[i]; L <= maxlevel; L++) {
            if (L > level[i])
                for (int k = 0; k < 6; k++) b[k] = (int)floorf((float)b[k] / 2.0f); /* arithmetic >> 1 */
            for (int x = b[0]; x <= b[1]; x++)
                for (int y = b[2]; y <= b[3]; y++)
                    for (int z = b[4]; z <= b[5]; z++) {
                        size_t h = hkey(L, x, y, z) & (hsize - 1);
                        for (hnode *nd = table[h]; nd; nd = nd->next) {
                            if (nd->level != L || nd->c[0] != x || nd->c[1] != y || nd->c[2] != z) continue;
                            int j = nd->geom;
                            if (j == i) continue;
                            if (L == level[i] && j < i) continue; /* same level: found from both sides */
                            /* lowest shared cell at level L */
                            int lo[3];
                            for (int k = 0; k < 3; k++) lo[k] = b[2 * k] > db[j][2 * k] ? b[2 * k] : db[j][2 * k];
                            if (lo[0] != x || lo[1] != y || lo[2] != z) continue;
                            if (pair_passes(&w->g[i], &w->g[j])) pb_push(pb, i, j);
                        }
                    }
        }
    }
    for (int bi = 0; bi < nbig; bi++) {
        int i = big[bi];
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            if (level[j] == 1000 && j < i) continue;
            if (pair_passes(&w->g[i], &w->g[j])) pb_push(pb, i, j);
        }
    }
    free(table); free(pool); free(level); free(db); free(big);
}

static void collect_pairs(orc_world *w, int method, pairbuf *pb) {
    for (int i = 0; i < w->ng; i++) geom_aabb(w, &w->g[i]);
    if (method == 1) {
        for (int i = 0; i < w->ng; i++)
            for (int j = i + 1; j < w->ng; j++)
                if (pair_passes(&w->g[i], &w->g[j])) pb_push(pb, i, j);
    } else {
        hash_broadphase(w, pb);
    }
    qsort(pb->p, (size_t)pb->n, sizeof(int) * 2, cmp_pair);
}

long orc_broadphase(orc_world *w, int method, int *pairs, long cap) {
    pairbuf pb = {0};
    collect_pairs(w, method, &pb);
    long nw = pb.n < cap ? pb.n : cap;
    if (pairs && nw > 0) memcpy(pairs, pb.p, sizeof(int) * 2 * (size_t)nw);
    long n = pb.n;
    free(pb.p);
    return n;
}

/* ------------------------------------------------------------------ colliders */

/* dCollideSpheres (sphere.cpp) */
static int collide_sphere_sphere(const ogeom *s1, const ogeom *s2, orc_contact_geom *c) {
    const float *p1 = s1->pos, *p2 = s2->pos;
    float r1 = s1->dims[0], r2 = s2->dims[0];
    float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
    float d = sqrtf(dx * dx + dy * dy + dz * dz);
    if (d > (r1 + r2)) return 0;
    if (d <= 0) {
        c->pos[0] = p1[0]; c->pos[1] = p1[1]; c->pos[2] = p1[2];
        c->normal[0] = 1; c->normal[1] = 0; c->normal[2] = 0;
        c->depth = r1 + r2;
    } else {
        float d1 = 1.0f / d;
        c->normal[0] = dx * d1; c->normal[1] = dy * d1; c->normal[2] = dz * d1;
        float k = 0.5f * (r2 - r1 - d);
        c->pos[0] = p1[0] + c->normal[0] * k;
        c->pos[1] = p1[1] + c->normal[1] * k;
        c->pos[2] = p1[2] + c->normal[2] * k;
        c->depth = r1 + r2 - d;
    }
    c->side1 = c->side2 = -1;
    return 1;
}

/* dCollideSphereBox (sphere.cpp) */
static int collide_sphere_box(const ogeom *s, const ogeom *bx, orc_contact_geom *c) {
    float l[3], t[3], p[3], q[3], r[3];
    int onborder = 0;
    const float *R = bx->R;
    for (int k = 0; k < 3; k++) p[k] = s->pos[k] - bx->pos[k];
    for (int k = 0; k < 3; k++) {
        l[k] = bx->dims[k] * 0.5f;
        t[k] = dot3_14(p, R + k);
        if (t[k] < -l[k]) { t[k] = -l[k]; onborder = 1; }
        if (t[k] > l[k]) { t[k] = l[k]; onborder = 1; }
    }
    c->side1 = c->side2 = -1;
    if (!onborder) {
        float min_distance = l[0] - fabsf(t[0]);
        int mini = 0;
        for (int i = 1; i < 3; i++) {
            float face_distance = l[i] - fabsf(t[i]);
            if (face_distance < min_distance) { min_distance = face_distance; mini = i; }
        }
        memcpy(c->pos, s->pos, 12);
        float tmp[3] = {0, 0, 0};
        tmp[mini] = (t[mini] > 0) ? 1.0f : -1.0f;
        mul0_331(c->normal, R, tmp);
        c->depth = min_distance + s->dims[0];
        return 1;
    }
    mul0_331(q, R, t);
    for (int k = 0; k < 3; k++) r[k] = p[k] - q[k];
    float depth = s->dims[0] - sqrtf(dot3(r, r));
    if (depth < 0) return 0;
    for (int k = 0; k < 3; k++) c->pos[k] = q[k] + bx->pos[k];
    memcpy(c->normal, r, 12);
    normalize3(c->normal);
    c->depth = depth;
    return 1;
}

/* dCollideSpherePlane (sphere.cpp) */
static int collide_sphere_plane(const ogeom *s, const ogeom *pl, orc_contact_geom *c) {
    const float *n = pl->dims;
    float k = dot3(s->pos, n);
    float depth = n[3] - k + s->dims[0];
    if (depth < 0) return 0;
    memcpy(c->normal, n, 12);
    for (int i = 0; i < 3; i++) c->pos[i] = s->pos[i] - n[i] * s->dims[0];
    c->depth = depth;
    c->side1 = c->side2 = -1;
    return 1;
}

/* dCollideBoxPlane (box.cpp), maxc capped at 4 */
static int collide_box_plane(const ogeom *bx, const ogeom *pl, int maxc, orc_contact_geom *c) {
    const float *R = bx->R, *n = pl->dims, *side = bx->dims;
    float Q[3], A[3], B[3];
    for (int k = 0; k < 3; k++) {
        Q[k] = dot3_14(n, R + k);
        A[k] = side[k] * Q[k];
        B[k] = fabsf(A[k]);
    }
    float depth = n[3] + 0.5f * (B[0] + B[1] + B[2]) - dot3(n, bx->pos);
    if (depth < 0) return 0;
    if (maxc > 4) maxc = 4;
    if (maxc < 1) maxc = 1;
    float p[3] = {bx->pos[0], bx->pos[1], bx->pos[2]};
    for (int i = 0; i < 3; i++) {
        float sgn = (A[i] > 0) ? -1.0f : 1.0f;
        /* p[k] -= / += 0.5*side[i]*R[k][i] */
        for (int k = 0; k < 3; k++) {
            float term = 0.5f * side[i] * R[k * 4 + i];
            if (sgn < 0) p[k] -= term; else p[k] += term;
        }
    }
    int ret = 1;
    memcpy(c[0].pos, p, 12);
    c[0].depth = depth;
    if (maxc > 1) {
        /* second and third contacts: walk along the two sides with the smallest projection */
        int first, second;
        if (B[0] < B[1]) {
            if (B[2] < B[0]) { first = 2; second = (B[0] < B[1]) ? 0 : 1; }
            else { first = 0; second = (B[1] < B[2]) ? 1 : 2; }
        } else {
            if (B[2] < B[1]) { first = 2; second = (B[0] < B[1]) ? 0 : 1; }
            else { first = 1; second = (B[0] < B[2]) ? 0 : 2; }
        }
        int order[2] = {first, second};
        for (int s = 0; s < 2; s++) {
            if (s == 1 && maxc == 2) break;
            int j = order[s];
            if (depth - B[j] < 0) break;
            for (int k = 0; k < 3; k++) {
                float term = side[j] * R[k * 4 + j];
                c[ret].pos[k] = (A[j] > 0) ? (p[k] + term) : (p[k] - term);
            }
            c[ret].depth = depth - B[j];
            ret++;
        }
    }
    if (maxc == 4 && ret == 3) {
        float d4 = c[1].depth + c[2].depth - depth;
        if (d4 > 0) {
            for (int k = 0; k < 3; k++) c[3].pos[k] = c[1].pos[k] + c[2].pos[k] - p[k];
            c[3].depth = d4;
            ret++;
        }
    }
    for (int i = 0; i < ret; i++) {
        memcpy(c[i].normal, n, 12);
        c[i].side1 = c[i].side2 = -1;
    }
    return ret;
}

/* intersectRectQuad (box.cpp): clip quad p (4 xy points) to the rectangle +-h; <= 8 points */
static int intersect_rect_quad(const float h[2], const float p[8], float ret[16]) {
    int nq = 4, nr = 0;
    float buffer[16];
    const float *q = p;
    float *r = ret;
    for (int dir = 0; dir <= 1; dir++) {
        for (int sign = -1; sign <= 1; sign += 2) {
            const float *pq = q;
            float *pr = r;
            nr = 0;
            for (int i = nq; i > 0; i--) {
                if (sign * pq[dir] < h[dir]) {
                    pr[0] = pq[0]; pr[1] = pq[1];
                    pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                const float *nextq = (i > 1) ? pq + 2 : q;
                if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
                    pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
                    pr[dir] = sign * h[dir];
                    pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                pq += 2;
            }
            q = r;
            r = (q == ret) ? buffer : ret;
            nq = nr;
        }
    }
done:
    if (q != ret) memcpy(ret, q, (size_t)nr * 2 * sizeof(float));
    return nr;
}

/* cullPoints (box.cpp): pick m of n points spread in angle around the centroid, i0 first */
static void cull_points(int n, const float p[], int m, int i0, int iret[]) {
    float a, cx, cy, q;
    if (n == 1) { cx = p[0]; cy = p[1]; }
    else if (n == 2) { cx = 0.5f * (p[0] + p[2]); cy = 0.5f * (p[1] + p[3]); }
    else {
        a = 0; cx = 0; cy = 0;
        for (int i = 0; i < n - 1; i++) {
            q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
            a += q;
            cx += q * (p[i * 2] + p[i * 2 + 2]);
            cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
        }
        q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
        a = 1.0f / (3.0f * (a + q));
        cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
        cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
    }
    float A[8];
    for (int i = 0; i < n; i++) A[i] = atan2f(p[i * 2 + 1] - cy, p[i * 2] - cx);
    int avail[8];
    for (int i = 0; i < n; i++) avail[i] = 1;
    avail[i0] = 0;
    iret[0] = i0;
    iret++;
    const float pi = 3.14159265358979323846f;
    for (int j = 1; j < m; j++) {
        a = (float)j * (2 * pi / m) + A[i0];
        if (a > pi) a -= 2 * pi;
        float maxdiff = 1e9f, diff;
        *iret = i0;
        for (int i = 0; i < n; i++) {
            if (avail[i]) {
                diff = fabsf(A[i] - a);
                if (diff > pi) diff = 2 * pi - diff;
                if (diff < maxdiff) { maxdiff = diff; *iret = i; }
            }
        }
        avail[*iret] = 0;
        iret++;
    }
}

/* dLineClosestApproach (box.cpp) */
static void line_closest_approach(const float pa[3], const float ua[3], const float pb[3], const float ub[3],
                                  float *alpha, float *beta) {
    float p[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    float uaub = dot3(ua, ub);
    float q1 = dot3(ua, p);
    float q2 = -dot3(ub, p);
    float d = 1 - uaub * uaub;
    if (d <= 0.0001f) { *alpha = 0; *beta = 0; }
    else {
        d = 1.0f / d;
        *alpha = (q1 + uaub * q2) * d;
        *beta = (uaub * q1 + q2) * d;
    }
}

/* dBoxBox (box.cpp): 15-axis SAT, then edge-edge point or face clipping. normal/depth/code out;
 * contacts written with pos+depth only. Returns the contact count. */
static int box_box(const float p1[3], const float R1[12], const float side1[3], const float p2[3],
                   const float R2[12], const float side2[3], float normal[3], float *depth_out,
                   int *code_out, int maxc_in, orc_contact_geom *contact) {
    const float fudge_factor = 1.05f;
    float p[3], pp[3], normalC[3] = {0, 0, 0};
    const float *normalR = 0;
    float A[3], B[3], Rm[3][3], Q[3][3], s, s2, l, e1;
    int invert_normal, code;

    for (int k = 0; k < 3; k++) p[k] = p2[k] - p1[k];
    mul1_331(pp, R1, p);
    for (int k = 0; k < 3; k++) { A[k] = side1[k] * 0.5f; B[k] = side2[k] * 0.5f; }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { Rm[i][j] = dot3_44(R1 + i, R2 + j); Q[i][j] = fabsf(Rm[i][j]); }

    s = -ORC_INF; invert_normal = 0; code = 0;

#define TST1(expr1, expr2, norm, cc)                 \
    e1 = (expr1);                                    \
    s2 = fabsf(e1) - (expr2);                        \
    if (s2 > 0) return 0;                            \
    if (s2 > s) { s = s2; normalR = (norm); invert_normal = (e1 < 0); code = (cc); }

    TST1(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), R1 + 0, 1)
    TST1(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), R1 + 1, 2)
    TST1(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), R1 + 2, 3)
    TST1(dot3_41(R2 + 0, p), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), R2 + 0, 4)
    TST1(dot3_41(R2 + 1, p), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), R2 + 1, 5)
    TST1(dot3_41(R2 + 2, p), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), R2 + 2, 6)
#undef TST1

#define TST2(expr1, expr2, n1, n2, n3, cc)                                  \
    e1 = (expr1);                                                           \
    s2 = fabsf(e1) - (expr2);                                               \
    if (s2 > 0) return 0;                                                   \
    l = sqrtf((n1) * (n1) + (n2) * (n2) + (n3) * (n3));                     \
    if (l > 0) {                                                            \
        s2 /= l;                                                            \
        if (s2 * fudge_factor > s) {                                        \
            s = s2; normalR = 0;                                            \
            normalC[0] = (n1) / l; normalC[1] = (n2) / l; normalC[2] = (n3) / l; \
            invert_normal = (e1 < 0); code = (cc);                          \
        }                                                                   \
    }

    TST2(pp[2] * Rm[1][0] - pp[1] * Rm[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -Rm[2][0], Rm[1][0], 7)
    TST2(pp[2] * Rm[1][1] - pp[1] * Rm[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -Rm[2][1], Rm[1][1], 8)
    TST2(pp[2] * Rm[1][2] - pp[1] * Rm[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -Rm[2][2], Rm[1][2], 9)
    TST2(pp[0] * Rm[2][0] - pp[2] * Rm[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), Rm[2][0], 0, -Rm[0][0], 10)
    TST2(pp[0] * Rm[2][1] - pp[2] * Rm[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), Rm[2][1], 0, -Rm[0][1], 11)
    TST2(pp[0] * Rm[2][2] - pp[2] * Rm[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), Rm[2][2], 0, -Rm[0][2], 12)
    TST2(pp[1] * Rm[0][0] - pp[0] * Rm[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -Rm[1][0], Rm[0][0], 0, 13)
    TST2(pp[1] * Rm[0][1] - pp[0] * Rm[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -Rm[1][1], Rm[0][1], 0, 14)
    TST2(pp[1] * Rm[0][2] - pp[0] * Rm[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -Rm[1][2], Rm[0][2], 0, 15)
#undef TST2

    if (!code) return 0;

    if (normalR) { normal[0] = normalR[0]; normal[1] = normalR[4]; normal[2] = normalR[8]; }
    else mul0_331(normal, R1, normalC);
    if (invert_normal) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
    *depth_out = -s;

    if (code > 6) {
        /* edge-edge: one contact at the midpoint of the closest points */
        float pa[3], pb[3], sign;
        for (int i = 0; i < 3; i++) pa[i] = p1[i];
        for (int j = 0; j < 3; j++) {
            sign = (dot3_14(normal, R1 + j) > 0) ? 1.0f : -1.0f;
            for (int i = 0; i < 3; i++) pa[i] += sign * A[j] * R1[i * 4 + j];
        }
        for (int i = 0; i < 3; i++) pb[i] = p2[i];
        for (int j = 0; j < 3; j++) {
            sign = (dot3_14(normal, R2 + j) > 0) ? -1.0f : 1.0f;
            for (int i = 0; i < 3; i++) pb[i] += sign * B[j] * R2[i * 4 + j];
        }
        float alpha, beta, ua[3], ub[3];
        for (int i = 0; i < 3; i++) ua[i] = R1[((code)-7) / 3 + i * 4];
        for (int i = 0; i < 3; i++) ub[i] = R2[((code)-7) % 3 + i * 4];
        line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
        for (int i = 0; i < 3; i++) pa[i] += ua[i] * alpha;
        for (int i = 0; i < 3; i++) pb[i] += ub[i] * beta;
        for (int i = 0; i < 3; i++) contact[0].pos[i] = 0.5f * (pa[i] + pb[i]);
        contact[0].depth = *depth_out;
        *code_out = code;
        return 1;
    }

    /* face-something: reference face 'a', incident face 'b' */
    const float *Ra, *Rb, *pa, *pb, *Sa, *Sb;
    if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; }
    else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; }

    float normal2[3], nr[3], anr[3];
    if (code <= 3) { normal2[0] = normal[0]; normal2[1] = normal[1]; normal2[2] = normal[2]; }
    else { normal2[0] = -normal[0]; normal2[1] = -normal[1]; normal2[2] = -normal[2]; }
    mul1_331(nr, Rb, normal2);
    anr[0] = fabsf(nr[0]); anr[1] = fabsf(nr[1]); anr[2] = fabsf(nr[2]);

    int lanr, a1, a2;
    if (anr[1] > anr[0]) {
        if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; }
        else { a1 = 0; a2 = 1; lanr = 2; }
    } else {
        if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; }
        else { a1 = 0; a2 = 1; lanr = 2; }
    }

    float center[3];
    if (nr[lanr] < 0) { for (int i = 0; i < 3; i++) center[i] = pb[i] - pa[i] + Sb[lanr] * Rb[i * 4 + lanr]; }
    else { for (int i = 0; i < 3; i++) center[i] = pb[i] - pa[i] - Sb[lanr] * Rb[i * 4 + lanr]; }

    int codeN, code1, code2;
    codeN = (code <= 3) ? code - 1 : code - 4;
    if (codeN == 0) { code1 = 1; code2 = 2; }
    else if (codeN == 1) { code1 = 0; code2 = 2; }
    else { code1 = 0; code2 = 1; }

    float quad[8], c1, c2, m11, m12, m21, m22;
    c1 = dot3_14(center, Ra + code1);
    c2 = dot3_14(center, Ra + code2);
    m11 = dot3_44(Ra + code1, Rb + a1);
    m12 = dot3_44(Ra + code1, Rb + a2);
    m21 = dot3_44(Ra + code2, Rb + a1);
    m22 = dot3_44(Ra + code2, Rb + a2);
    {
        float k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
        quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4;
        quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
        quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4;
        quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
    }
    float rect[2] = {Sa[code1], Sa[code2]};
    float ret[16];
    int n = intersect_rect_quad(rect, quad, ret);
    if (n < 1) return 0;

    float point[3 * 8], dep[8];
    float det1 = 1.0f / (m11 * m22 - m12 * m21);
    m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
    int cnum = 0;
    for (int j = 0; j < n; j++) {
        float k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
        float k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
        for (int i = 0; i < 3; i++) point[cnum * 3 + i] = center[i] + k1 * Rb[i * 4 + a1] + k2 * Rb[i * 4 + a2];
        dep[cnum] = Sa[codeN] - dot3(normal2, point + cnum * 3);
        if (dep[cnum] >= 0) {
            ret[cnum * 2] = ret[j * 2];
            ret[cnum * 2 + 1] = ret[j * 2 + 1];
            cnum++;
        }
    }
    if (cnum < 1) return 0;

    int maxc = maxc_in;
    if (maxc > cnum) maxc = cnum;
    if (maxc < 1) maxc = 1;

    if (cnum <= maxc) {
        if (code < 4) {
            for (int j = 0; j < cnum; j++) {
                for (int i = 0; i < 3; i++) contact[j].pos[i] = point[j * 3 + i] + pa[i];
                contact[j].depth = dep[j];
            }
        } else {
            for (int j = 0; j < cnum; j++) {
                for (int i = 0; i < 3; i++) contact[j].pos[i] = point[j * 3 + i] + pa[i] - normal[i] * dep[j];
                contact[j].depth = dep[j];
            }
        }
    } else {
        int i1 = 0;
        float maxdepth = dep[0];
        for (int i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
        int iret[8];
        cull_points(cnum, ret, maxc, i1, iret);
        for (int j = 0; j < maxc; j++) {
            for (int i = 0; i < 3; i++) contact[j].pos[i] = point[iret[j] * 3 + i] + pa[i];
            contact[j].depth = dep[iret[j]];
        }
        cnum = maxc;
    }
    *code_out = code;
    return cnum;
}

/* dCollideBoxBox (box.cpp): contact normal = -dBoxBox normal */
static int collide_box_box(const ogeom *b1, const ogeom *b2, int maxc, orc_contact_geom *c) {
    float normal[3], depth;
    int code;
    if (maxc > 8) maxc = 8;
    int num = box_box(b1->pos, b1->R, b1->dims, b2->pos, b2->R, b2->dims, normal, &depth, &code, maxc, c);
    for (int i = 0; i < num; i++) {
        c[i].normal[0] = -normal[0]; c[i].normal[1] = -normal[1]; c[i].normal[2] = -normal[2];
        c[i].side1 = c[i].side2 = -1;
    }
    return num;
}

/* closest point on triangle abc to p (Ericson, Real-Time Collision Detection 5.1.5) */
static void closest_pt_triangle(const float p[3], const float a[3], const float b[3], const float c[3], float out[3]) {
    float ab[3], ac[3], ap[3], bp[3], cp[3];
    for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) { memcpy(out, a, 12); return; }
    for (int k = 0; k < 3; k++) bp[k] = p[k] - b[k];
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) { memcpy(out, b, 12); return; }
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        float v = d1 / (d1 - d3);
        for (int k = 0; k < 3; k++) out[k] = a[k] + v * ab[k];
        return;
    }
    for (int k = 0; k < 3; k++) cp[k] = p[k] - c[k];
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) { memcpy(out, c, 12); return; }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        float wv = d2 / (d2 - d6);
        for (int k = 0; k < 3; k++) out[k] = a[k] + wv * ac[k];
        return;
    }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        float wv = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; k++) out[k] = b[k] + wv * (c[k] - b[k]);
        return;
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, wv = vc * denom;
    for (int k = 0; k < 3; k++) out[k] = a[k] + ab[k] * v + ac[k] * wv;
}

/* sphere vs trimesh. libode's dCollideSTL is the least certain item of SURVEY.md A.4, so the
 * engine defines its own order-independent rule, restated here:
 *   per triangle (mesh-local frame): skip if the sphere's box misses the triangle's box; q =
 *   closest point; dist = |c - q|; candidate iff dist <= r; dir = (c-q)/dist, or the triangle
 *   normal when dist == 0; depth = r - dist.
 *   selection: order candidates by (depth desc, triangle index asc); accept greedily unless an
 *   accepted contact lies within 1e-3*r of it; stop at maxc (<= 8).
 * Output (sphere = g1, trimesh = g2): pos = q, normal = dir (from the mesh into the sphere),
 * side2 = triangle index. */
typedef struct { float depth; int tri; float q[3], n[3]; } tri_cand;
static int cmp_cand(const void *x, const void *y) {
    const tri_cand *a = (const tri_cand *)x, *b = (const tri_cand *)y;
    if (a->depth != b->depth) return a->depth > b->depth ? -1 : 1;
    return a->tri < b->tri ? -1 : (a->tri > b->tri ? 1 : 0);
}
static int collide_sphere_trimesh(const orc_world *w, const ogeom *s, const ogeom *tm, int maxc, orc_contact_geom *out) {
    const omesh *m = &w->m[tm->mesh];
    float r = s->dims[0], c[3], d[3];
    for (int k = 0; k < 3; k++) d[k] = s->pos[k] - tm->pos[k];
    mul1_331(c, tm->R, d); /* sphere centre in the mesh frame */
    if (maxc > 8) maxc = 8;
    tri_cand *cand = 0;
    int nc = 0, capc = 0;
    for (int t = 0; t < m->nt; t++) {
        const float *a = m->v + 3 * m->t[3 * t], *b = m->v + 3 * m->t[3 * t + 1], *cc = m->v + 3 * m->t[3 * t + 2];
        int skip = 0;
        for (int k = 0; k < 3; k++) {
            float lo = fminf(a[k], fminf(b[k], cc[k])), hi = fmaxf(a[k], fmaxf(b[k], cc[k]));
            if (c[k] - r > hi || c[k] + r < lo) skip = 1;
        }
        if (skip) continue;
        float q[3], dv[3];
        closest_pt_triangle(c, a, b, cc, q);
        for (int k = 0; k < 3; k++) dv[k] = c[k] - q[k];
        float d2 = dot3(dv, dv);
        if (d2 > r * r) continue;
        float dist = sqrtf(d2);
        if (dist > r) continue;
        float n[3];
        if (dist > 0) {
            float inv = 1.0f / dist;
            n[0] = dv[0] * inv; n[1] = dv[1] * inv; n[2] = dv[2] * inv;
        } else {
            float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {cc[0] - a[0], cc[1] - a[1], cc[2] - a[2]};
            cross3(n, e1, e2);
            float l2 = dot3(n, n);
            if (!(l2 > 0)) continue;
            float inv = 1.0f / sqrtf(l2);
            n[0] *= inv; n[1] *= inv; n[2] *= inv;
        }
        if (nc == capc) { capc = capc ? capc * 2 : 32; cand = (tri_cand *)realloc(cand, sizeof(tri_cand) * (size_t)capc); }
        cand[nc].depth = r - dist; cand[nc].tri = t;
        memcpy(cand[nc].q, q, 12); memcpy(cand[nc].n, n, 12);
        nc++;
    }
    qsort(cand, (size_t)nc, sizeof(tri_cand), cmp_cand);
    int nout = 0;
    float tol2 = (1e-3f * r) * (1e-3f * r);
    float accq[8][3];
    for (int i = 0; i < nc && nout < maxc; i++) {
        int dup = 0;
        for (int j = 0; j < nout; j++) {
            float e[3] = {cand[i].q[0] - accq[j][0], cand[i].q[1] - accq[j][1], cand[i].q[2] - accq[j][2]};
            if (dot3(e, e) <= tol2) { dup = 1; break; }
        }
        if (dup) continue;
        memcpy(accq[nout], cand[i].q, 12);
        float pw[3], nw[3];
        mul0_331(pw, tm->R, cand[i].q);
        mul0_331(nw, tm->R, cand[i].n);
        for (int k = 0; k < 3; k++) { out[nout].pos[k] = pw[k] + tm->pos[k]; out[nout].normal[k] = nw[k]; }
        out[nout].depth = cand[i].depth;
        out[nout].side1 = -1; out[nout].side2 = cand[i].tri;
        nout++;
    }
    free(cand);
    return nout;
}

/* box vs trimesh.  libode's collision_trimesh_box.cpp (SAT + clipping per triangle) is not restated; like
 * sphere-trimesh the engine defines an order-independent vertex/face rule, restated here (mesh-local frame):
 *   per triangle with unit normal n (skipped when degenerate or when its box misses the box's AABB):
 *     (a) each of the 8 box vertices pv with signed distance d = n.(pv - a) in [-maxdepth, 0) whose
 *         projection lies inside the triangle (three edge tests, boundary included): contact at pv,
 *         direction n, depth -d; maxdepth = the box's smallest side length;
 *     (b) each triangle vertex inside the box: contact at the vertex, direction = minus the outward normal
 *         of the box face it is nearest to (lowest axis on ties), depth = distance to that face.
 *   selection as for spheres: (depth desc, triangle*16 + sub-index asc), duplicates within 1e-3 * the
 *   smallest half side dropped, at most maxc (<= 8).  Edge-edge contacts are not generated.
 * Output (box = g1, trimesh = g2): normal from the mesh into the box, side2 = triangle index. */
static int collide_box_trimesh(const orc_world *w, const ogeom *bx, const ogeom *tm, int maxc, orc_contact_geom *out) {
    const omesh *m = &w->m[tm->mesh];
    float cb[3], d[3], A[3][3], h[3], ext[3], pv[8][3];
    for (int k = 0; k < 3; k++) d[k] = bx->pos[k] - tm->pos[k];
    mul1_331(cb, tm->R, d);
    for (int k = 0; k < 3; k++) {
        float col[3] = {bx->R[k], bx->R[4 + k], bx->R[8 + k]};
        mul1_331(A[k], tm->R, col);
        h[k] = 0.5f * bx->dims[k];
    }
    for (int j = 0; j < 3; j++) ext[j] = fabsf(A[0][j]) * h[0] + fabsf(A[1][j]) * h[1] + fabsf(A[2][j]) * h[2];
    for (int v = 0; v < 8; v++)
        for (int j = 0; j < 3; j++) {
            float s0 = (v & 1) ? h[0] : -h[0], s1 = (v & 2) ? h[1] : -h[1], s2 = (v & 4) ? h[2] : -h[2];
            pv[v][j] = ((cb[j] + s0 * A[0][j]) + s1 * A[1][j]) + s2 * A[2][j];
        }
    const float hmin = fminf(h[0], fminf(h[1], h[2]));
    const float maxdepth = 2.0f * hmin;
    if (maxc > 8) maxc = 8;
    tri_cand *cand = 0;
    int nc = 0, capc = 0;
    for (int t = 0; t < m->nt; t++) {
        const float *a = m->v + 3 * m->t[3 * t], *b = m->v + 3 * m->t[3 * t + 1], *c = m->v + 3 * m->t[3 * t + 2];
        int skip = 0;
        for (int k = 0; k < 3; k++) {
            float lo = fminf(a[k], fminf(b[k], c[k])), hi = fmaxf(a[k], fmaxf(b[k], c[k]));
            if (cb[k] - ext[k] > hi || cb[k] + ext[k] < lo) skip = 1;
        }
        if (skip) continue;
        float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nr[3], n[3];
        cross3(nr, e1, e2);
        float l2 = dot3(nr, nr);
        if (!(l2 > 0)) continue;
        float inv = 1.0f / sqrtf(l2);
        n[0] = nr[0] * inv; n[1] = nr[1] * inv; n[2] = nr[2] * inv;
        for (int sub = 0; sub < 11; sub++) {
            float depth, q[3], dir[3];
            if (sub < 8) {
                float e[3] = {pv[sub][0] - a[0], pv[sub][1] - a[1], pv[sub][2] - a[2]};
                float dd = dot3(n, e);
                if (!(dd < 0 && dd >= -maxdepth)) continue;
                float pr[3] = {pv[sub][0] - dd * n[0], pv[sub][1] - dd * n[1], pv[sub][2] - dd * n[2]};
                const float *tv[3] = {a, b, c};
                int inside = 1;
                for (int j = 0; j < 3; j++) {
                    const float *p0 = tv[j], *p1 = tv[(j + 1) % 3];
                    float ed[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, rq[3] = {pr[0] - p0[0], pr[1] - p0[1], pr[2] - p0[2]}, cr[3];
                    cross3(cr, ed, rq);
                    if (!(dot3(cr, nr) >= 0)) inside = 0;
                }
                if (!inside) continue;
                depth = -dd;
                memcpy(q, pv[sub], 12); memcpy(dir, n, 12);
            } else {
                const float *tv = sub == 8 ? a : (sub == 9 ? b : c);
                float e[3] = {tv[0] - cb[0], tv[1] - cb[1], tv[2] - cb[2]}, loc[3];
                int inside = 1, ks = 0;
                float best = 0;
                for (int k = 0; k < 3; k++) {
                    loc[k] = dot3(A[k], e);
                    if (!(fabsf(loc[k]) <= h[k])) inside = 0;
                    float pen = h[k] - fabsf(loc[k]);
                    if (k == 0 || pen < best) { best = pen; ks = k; }
                }
                if (!inside) continue;
                float sg = loc[ks] < 0 ? 1.0f : -1.0f; /* minus the outward normal of the nearest face */
                depth = best;
                memcpy(q, tv, 12);
                dir[0] = sg * A[ks][0]; dir[1] = sg * A[ks][1]; dir[2] = sg * A[ks][2];
            }
            if (nc == capc) { capc = capc ? capc * 2 : 32; cand = (tri_cand *)realloc(cand, sizeof(tri_cand) * (size_t)capc); }
            cand[nc].depth = depth; cand[nc].tri = t * 16 + sub;
            memcpy(cand[nc].q, q, 12); memcpy(cand[nc].n, dir, 12);
            nc++;
        }
    }
    qsort(cand, (size_t)nc, sizeof(tri_cand), cmp_cand);
    int nout = 0;
    float tol2 = (1e-3f * hmin) * (1e-3f * hmin);
    float accq[8][3];
    for (int i = 0; i < nc && nout < maxc; i++) {
        int dup = 0;
        for (int j = 0; j < nout; j++) {
            float e[3] = {cand[i].q[0] - accq[j][0], cand[i].q[1] - accq[j][1], cand[i].q[2] - accq[j][2]};
            if (dot3(e, e) <= tol2) { dup = 1; break; }
        }
        if (dup) continue;
        memcpy(accq[nout], cand[i].q, 12);
        float pw[3], nw[3];
        mul0_331(pw, tm->R, cand[i].q);
        mul0_331(nw, tm->R, cand[i].n);
        for (int k = 0; k < 3; k++) { out[nout].pos[k] = pw[k] + tm->pos[k]; out[nout].normal[k] = nw[k]; }
        out[nout].depth = cand[i].depth;
        out[nout].side1 = -1; out[nout].side2 = cand[i].tri >> 4;
        nout++;
    }
    free(cand);
    return nout;
}

/* dCollide (collision_kernel.cpp): colliders are registered for (lower class, higher class);
 * called the other way round, the result is computed swapped and then normals are negated and
 * g1/g2, side1/side2 exchanged. */
int orc_collide(orc_world *w, int i1, int i2, int maxc, orc_contact_geom *out) {
    ogeom *g1 = &w->g[i1], *g2 = &w->g[i2];
    if (i1 == i2 || maxc < 1) return 0;
    if (g1->body == g2->body && g1->body >= 0) return 0;
    geom_sync_pose(w, g1);
    geom_sync_pose(w, g2);
    int swap = g1->type > g2->type;
    ogeom *a = swap ? g2 : g1, *b = swap ? g1 : g2;
    int n = 0;
    if (a->type == ORC_SPHERE && b->type == ORC_SPHERE) n = collide_sphere_sphere(a, b, out);
    else if (a->type == ORC_SPHERE && b->type == ORC_BOX) n = collide_sphere_box(a, b, out);
    else if (a->type == ORC_SPHERE && b->type == ORC_PLANE) n = collide_sphere_plane(a, b, out);
    else if (a->type == ORC_BOX && b->type == ORC_BOX) n = collide_box_box(a, b, maxc, out);
    else if (a->type == ORC_BOX && b->type == ORC_PLANE) n = collide_box_plane(a, b, maxc, out);
    else if (a->type == ORC_SPHERE && b->type == ORC_TRIMESH) n = collide_sphere_trimesh(w, a, b, maxc, out);
    else if (a->type == ORC_BOX && b->type == ORC_TRIMESH) n = collide_box_trimesh(w, a, b, maxc, out);
    else n = 0; /* plane-plane, plane-trimesh, ...: no collider */
    for (int i = 0; i < n; i++) {
        if (swap) {
            out[i].normal[0] = -out[i].normal[0]; out[i].normal[1] = -out[i].normal[1]; out[i].normal[2] = -out[i].normal[2];
            int t = out[i].side1; out[i].side1 = out[i].side2; out[i].side2 = t;
        }
        out[i].g1 = i1; out[i].g2 = i2;
    }
    return n;
}

/* ------------------------------------------------------------------ contact joints */

int orc_add_contact(orc_world *w, const orc_contact_geom *g, const orc_surface *s, int b1, int b2) {
    if (w->nj == w->capj) {
        w->capj = w->capj ? w->capj * 2 : 256;
        w->j = (ojoint *)realloc(w->j, sizeof(ojoint) * (size_t)w->capj);
    }
    ojoint *j = &w->j[w->nj];
    j->g = *g; j->s = *s;
    /* dJointAttach: a NULL body1 is swapped into slot 2 and the joint is flagged REVERSE */
    if (b1 < 0 && b2 >= 0) { j->b1 = b2; j->b2 = -1; j->reverse = 1; }
    else { j->b1 = b1; j->b2 = b2; j->reverse = 0; }
    return w->nj++;
}
void orc_clear_contacts(orc_world *w) { w->nj = 0; }
int orc_num_contacts(const orc_world *w) { return w->nj; }

long orc_collide_all(orc_world *w, int maxc, const orc_surface *surf) {
    pairbuf pb = {0};
    collect_pairs(w, 0, &pb);
    long total = 0;
    orc_contact_geom cg[8];
    if (maxc > 8) maxc = 8;
    for (long i = 0; i < pb.n; i++) {
        int g1 = pb.p[2 * i], g2 = pb.p[2 * i + 1];
        /* the engine's canonical callback order: lower class first, then lower id */
        if (w->g[g1].type > w->g[g2].type) { int t = g1; g1 = g2; g2 = t; }
        int n = orc_collide(w, g1, g2, maxc, cg);
        for (int k = 0; k < n; k++) orc_add_contact(w, &cg[k], surf, w->g[g1].body, w->g[g2].body);
        total += n;
    }
    free(pb.p);
    return total;
}

/* ------------------------------------------------------------------ QuickStep */

/* ODE's dRand / dRandInt (misc.cpp) */
static unsigned long ode_rand(orc_world *w) {
    w->lcg_seed = (1664525ul * w->lcg_seed + 1013904223ul) & 0xfffffffful;
    return w->lcg_seed;
}
static int ode_rand_int(orc_world *w, int n) {
    unsigned long un = (unsigned long)n;
    unsigned long r = ode_rand(w);
    if (un <= 0x00010000ul) {
        r ^= (r >> 16);
        if (un <= 0x00000100ul) {
            r ^= (r >> 8);
            if (un <= 0x00000010ul) {
                r ^= (r >> 4);
                if (un <= 0x00000004ul) {
                    r ^= (r >> 2);
                    if (un <= 0x00000002ul) r ^= (r >> 1);
                }
            }
        }
    }
    return (int)(r % un);
}

static int joint_rows(const ojoint *j) {
    /* dxJointContact::getInfo1 */
    int m = 1;
    float mu = j->s.mu < 0 ? 0 : j->s.mu;
    if (j->s.mode & ORC_MU2) {
        float mu2 = j->s.mu2 < 0 ? 0 : j->s.mu2;
        if (mu > 0) m++;
        if (mu2 > 0) m++;
    } else {
        if (mu > 0) m += 2;
    }
    return m;
}

int orc_num_rows(const orc_world *w) { return w->nrows; }
void orc_set_keep_rows(orc_world *w, int on) { w->keep_rows = on; if (!on) dbg_free(w); }
/* rows of the last step as dxJointContact::getInfo2 + QuickStep built them, BEFORE SOR_LCP's Ad scaling:
 * J (12 per row), c (the row's target velocity), cfm (already divided by h), lo, hi, rhs, body pair */
int orc_last_rows(const orc_world *w, float *J, float *c, float *cfm, float *lo, float *hi, float *rhs, int *jb, int n) {
    if (n > w->dbg_m) n = w->dbg_m;
    if (n > 0) {
        if (J) memcpy(J, w->dbg_J, sizeof(float) * 12 * (size_t)n);
        if (c) memcpy(c, w->dbg_c, sizeof(float) * (size_t)n);
        if (cfm) memcpy(cfm, w->dbg_cfm, sizeof(float) * (size_t)n);
        if (lo) memcpy(lo, w->dbg_lo, sizeof(float) * (size_t)n);
        if (hi) memcpy(hi, w->dbg_hi, sizeof(float) * (size_t)n);
        if (rhs) memcpy(rhs, w->dbg_rhs, sizeof(float) * (size_t)n);
        if (jb) memcpy(jb, w->dbg_jb, sizeof(int) * 2 * (size_t)n);
    }
    return w->dbg_m;
}
void orc_last_lambda(const orc_world *w, float *lambda, int n) {
    if (n > w->nrows) n = w->nrows;
    if (n > 0) memcpy(lambda, w->last_lambda, sizeof(float) * (size_t)n);
}

/* Exact box-constrained LCP for an SPD matrix: block principal pivoting (Judice & Pires) with the
 * single-pivot fallback that guarantees termination.  F = free (w = 0), L = at lo (w >= 0), U = at hi
 * (w <= 0), w = A x - b.  Dense Cholesky on the free block; sizes here are a few hundred rows. */
static int solve_blcp(int m, const double *A, const double *b, const float *lo, const float *hi, double *x) {
    char *st = (char *)calloc((size_t)m, 1); /* 0 F, 1 L, 2 U */
    int *F = (int *)malloc(sizeof(int) * (size_t)m);
    double *C = (double *)malloc(sizeof(double) * (size_t)m * (size_t)m);
    double *r = (double *)malloc(sizeof(double) * (size_t)m);
    double *wv = (double *)malloc(sizeof(double) * (size_t)m);
    const double eps = 1e-11;
    int best = m + 1, p = 10, ok = 0;
    for (long iter = 0; iter < 200L * (m + 10); iter++) {
        int nf = 0;
        for (int i = 0; i < m; i++) {
            if (st[i] == 0) F[nf++] = i;
            else x[i] = st[i] == 1 ? (double)lo[i] : (double)hi[i];
        }
        for (int a = 0; a < nf; a++) {
            int i = F[a];
            double s = b[i];
            for (int j = 0; j < m; j++) if (st[j]) s -= A[(size_t)i * m + j] * x[j];
            r[a] = s;
            for (int c = 0; c <= a; c++) C[(size_t)a * nf + c] = A[(size_t)i * m + F[c]];
        }
        /* Cholesky C = L L^T (lower), then two triangular solves */
        for (int a = 0; a < nf; a++) {
            for (int c = 0; c <= a; c++) {
                double s = C[(size_t)a * nf + c];
                for (int k = 0; k < c; k++) s -= C[(size_t)a * nf + k] * C[(size_t)c * nf + k];
                if (a == c) { if (s <= 0) s = 1e-300; C[(size_t)a * nf + a] = sqrt(s); }
                else C[(size_t)a * nf + c] = s / C[(size_t)c * nf + c];
            }
        }
        for (int a = 0; a < nf; a++) {
            double s = r[a];
            for (int k = 0; k < a; k++) s -= C[(size_t)a * nf + k] * r[k];
            r[a] = s / C[(size_t)a * nf + a];
        }
        for (int a = nf - 1; a >= 0; a--) {
            double s = r[a];
            for (int k = a + 1; k < nf; k++) s -= C[(size_t)k * nf + a] * r[k];
            r[a] = s / C[(size_t)a * nf + a];
        }
        for (int a = 0; a < nf; a++) x[F[a]] = r[a];
        int ninf = 0, last = -1;
        for (int i = 0; i < m; i++) {
            double s = -b[i];
            for (int j = 0; j < m; j++) s += A[(size_t)i * m + j] * x[j];
            wv[i] = s;
            int bad = 0;
            if (st[i] == 0) bad = (x[i] < (double)lo[i] - eps) || (x[i] > (double)hi[i] + eps);
            else if (st[i] == 1) bad = wv[i] < -eps;
            else bad = wv[i] > eps;
            if (bad) { ninf++; last = i; }
        }
        if (ninf == 0) { ok = 1; break; }
        int all = 0;
        if (ninf < best) { best = ninf; p = 10; all = 1; }
        else if (p > 0) { p--; all = 1; }
        for (int i = 0; i < m; i++) {
            if (!all && i != last) continue;
            if (st[i] == 0) {
                if (x[i] < (double)lo[i] - eps) st[i] = 1;
                else if (x[i] > (double)hi[i] + eps) st[i] = 2;
            } else if ((st[i] == 1 && wv[i] < -eps) || (st[i] == 2 && wv[i] > eps)) st[i] = 0;
        }
    }
    free(st); free(F); free(C); free(r); free(wv);
    return ok;
}

int orc_quickstep(orc_world *w, float h, int order_mode, const int *perm) {
    int nb = w->nb;
    float stepsize1 = 1.0f / h;
    float *invI = (float *)malloc(sizeof(float) * 12 * (size_t)(nb ? nb : 1));

    /* per body: world-frame inverse inertia, gyroscopic torque, gravity */
    for (int i = 0; i < nb; i++) {
        obody *b = &w->b[i];
        float tmp[12];
        mul2_333(tmp, b->invI, b->R);
        mul0_333(invI + 12 * i, b->R, tmp);
        if ((b->flags & ORC_BODY_GYRO) && !(b->flags & ORC_BODY_KINEMATIC)) {
            float I[12], L[3], t[3];
            mul2_333(tmp, b->I, b->R);
            mul0_333(I, b->R, tmp);
            mul0_331(L, I, b->avel);
            cross3(t, b->avel, L);
            b->tacc[0] -= t[0]; b->tacc[1] -= t[1]; b->tacc[2] -= t[2];
        }
        if (!(b->flags & ORC_BODY_NOGRAVITY))
            for (int k = 0; k < 3; k++) b->facc[k] += b->mass * w->gravity[k];
    }

    /* rows. Joints whose both bodies are NULL are inert (never reached by island traversal). */
    int m = 0;
    int *jrow0 = (int *)malloc(sizeof(int) * (size_t)(w->nj + 1));
    for (int j = 0; j < w->nj; j++) {
        jrow0[j] = m;
        if (w->j[j].b1 >= 0) m += joint_rows(&w->j[j]);
    }
    jrow0[w->nj] = m;
    free(w->last_lambda);
    w->last_lambda = (float *)calloc((size_t)(m ? m : 1), sizeof(float));
    w->nrows = m;
    if (w->keep_rows && m == 0) dbg_free(w);

    if (m > 0) {
        float *J = (float *)calloc((size_t)m * 12, sizeof(float));
        float *iMJ = (float *)calloc((size_t)m * 12, sizeof(float));
        float *c = (float *)calloc((size_t)m, sizeof(float));
        float *cfm = (float *)malloc(sizeof(float) * (size_t)m);
        float *lo = (float *)malloc(sizeof(float) * (size_t)m);
        float *hi = (float *)malloc(sizeof(float) * (size_t)m);
        float *rhs = (float *)malloc(sizeof(float) * (size_t)m);
        float *Ad = (float *)malloc(sizeof(float) * (size_t)m);
        float *Adcfm = (float *)malloc(sizeof(float) * (size_t)m);
        int *findex = (int *)malloc(sizeof(int) * (size_t)m);
        int *jb = (int *)malloc(sizeof(int) * 2 * (size_t)m);
        float *lambda = w->last_lambda;

        for (int j = 0; j < w->nj; j++) {
            const ojoint *jt = &w->j[j];
            if (jt->b1 < 0) continue;
            int r0 = jrow0[j], the_m = jrow0[j + 1] - jrow0[j];
            const obody *B1 = &w->b[jt->b1];
            const obody *B2 = jt->b2 >= 0 ? &w->b[jt->b2] : 0;
            for (int r = 0; r < the_m; r++) {
                cfm[r0 + r] = w->cfm; lo[r0 + r] = -ORC_INF; hi[r0 + r] = ORC_INF; findex[r0 + r] = -1;
                jb[2 * (r0 + r)] = jt->b1; jb[2 * (r0 + r) + 1] = jt->b2;
            }
            /* dxJointContact::getInfo2 */
            float normal[3], c1[3], c2[3] = {0, 0, 0};
            for (int k = 0; k < 3; k++) normal[k] = jt->reverse ? -jt->g.normal[k] : jt->g.normal[k];
            for (int k = 0; k < 3; k++) c1[k] = jt->g.pos[k] - B1->pos[k];
            float *Jr = J + 12 * r0;
            memcpy(Jr, normal, 12);
            cross3(Jr + 3, c1, normal);
            if (B2) {
                for (int k = 0; k < 3; k++) c2[k] = jt->g.pos[k] - B2->pos[k];
                Jr[6] = -normal[0]; Jr[7] = -normal[1]; Jr[8] = -normal[2];
                cross3(Jr + 9, c2, normal);
                Jr[9] = -Jr[9]; Jr[10] = -Jr[10]; Jr[11] = -Jr[11];
            }
            float erp = w->erp;
            if (jt->s.mode & ORC_SOFT_ERP) erp = jt->s.soft_erp;
            float k = stepsize1 * erp;
            float depth = jt->g.depth - w->min_depth;
            if (depth < 0) depth = 0;
            if (jt->s.mode & ORC_SOFT_CFM) cfm[r0] = jt->s.soft_cfm;
            float motionN = 0;
            if (jt->s.mode & ORC_MOTIONN) motionN = jt->s.motionN;
            float pushout = k * depth + motionN;
            c[r0] = pushout;
            if (c[r0] > w->max_vel) c[r0] = w->max_vel;
            if (jt->s.mode & ORC_BOUNCE) {
                float outgoing = dot3(Jr, B1->lvel) + dot3(Jr + 3, B1->avel);
                if (B2) outgoing += dot3(Jr + 6, B2->lvel) + dot3(Jr + 9, B2->avel);
                outgoing -= motionN;
                if (jt->s.bounce_vel >= 0 && (-outgoing) > jt->s.bounce_vel) {
                    float newc = -jt->s.bounce * outgoing + motionN;
                    if (newc > c[r0]) c[r0] = newc;
                }
            }
            lo[r0] = 0; hi[r0] = ORC_INF;
            float t1[3], t2[3];
            if (the_m >= 2) {
                if (jt->s.mode & ORC_FDIR1) { memcpy(t1, jt->s.fdir1, 12); cross3(t2, normal, t1); }
                else orc_plane_space(normal, t1, t2);
                float *J1 = Jr + 12;
                memcpy(J1, t1, 12);
                cross3(J1 + 3, c1, t1);
                if (B2) {
                    J1[6] = -t1[0]; J1[7] = -t1[1]; J1[8] = -t1[2];
                    cross3(J1 + 9, c2, t1);
                    J1[9] = -J1[9]; J1[10] = -J1[10]; J1[11] = -J1[11];
                }
                if (jt->s.mode & ORC_MOTION1) c[r0 + 1] = jt->s.motion1;
                float mu = jt->s.mu < 0 ? 0 : jt->s.mu;
                lo[r0 + 1] = -mu; hi[r0 + 1] = mu;
                if (jt->s.mode & ORC_APPROX1_1) findex[r0 + 1] = r0;
                if (jt->s.mode & ORC_SLIP1) cfm[r0 + 1] = jt->s.slip1;
            }
            if (the_m >= 3) {
                float *J2 = Jr + 24;
                memcpy(J2, t2, 12);
                cross3(J2 + 3, c1, t2);
                if (B2) {
                    J2[6] = -t2[0]; J2[7] = -t2[1]; J2[8] = -t2[2];
                    cross3(J2 + 9, c2, t2);
                    J2[9] = -J2[9]; J2[10] = -J2[10]; J2[11] = -J2[11];
                }
                if (jt->s.mode & ORC_MOTION2) c[r0 + 2] = jt->s.motion2;
                float mu = jt->s.mu < 0 ? 0 : jt->s.mu;
                if (jt->s.mode & ORC_MU2) { float mu2 = jt->s.mu2 < 0 ? 0 : jt->s.mu2; lo[r0 + 2] = -mu2; hi[r0 + 2] = mu2; }
                else { lo[r0 + 2] = -mu; hi[r0 + 2] = mu; }
                if (jt->s.mode & ORC_APPROX1_2) findex[r0 + 2] = r0;
                if (jt->s.mode & ORC_SLIP2) cfm[r0 + 2] = jt->s.slip2;
            }
        }

        /* rhs = c/h - J (v/h + invM fe) ; cfm /= h */
        float *tmp1 = (float *)malloc(sizeof(float) * 6 * (size_t)nb);
        for (int i = 0; i < nb; i++) {
            const obody *b = &w->b[i];
            for (int k = 0; k < 3; k++) tmp1[6 * i + k] = b->facc[k] * b->invMass + b->lvel[k] * stepsize1;
            float t[3];
            mul0_331(t, invI + 12 * i, b->tacc);
            for (int k = 0; k < 3; k++) tmp1[6 * i + 3 + k] = t[k] + b->avel[k] * stepsize1;
        }
        for (int i = 0; i < m; i++) {
            int b1 = jb[2 * i], b2 = jb[2 * i + 1];
            float sum = 0;
            for (int k = 0; k < 6; k++) sum += J[12 * i + k] * tmp1[6 * b1 + k];
            if (b2 >= 0) for (int k = 0; k < 6; k++) sum += J[12 * i + 6 + k] * tmp1[6 * b2 + k];
            rhs[i] = c[i] * stepsize1 - sum;
            cfm[i] *= stepsize1;
        }
        free(tmp1);

        /* iMJ = invM * J^T (compute_invM_JT) */
        for (int i = 0; i < m; i++) {
            int b1 = jb[2 * i], b2 = jb[2 * i + 1];
            float *im = iMJ + 12 * i, *Ji = J + 12 * i;
            for (int k = 0; k < 3; k++) im[k] = w->b[b1].invMass * Ji[k];
            mul0_331(im + 3, invI + 12 * b1, Ji + 3);
            if (b2 >= 0) {
                for (int k = 0; k < 3; k++) im[6 + k] = w->b[b2].invMass * Ji[6 + k];
                mul0_331(im + 9, invI + 12 * b2, Ji + 9);
            }
        }
        if (w->keep_rows) {
            dbg_free(w);
            w->dbg_m = m;
            w->dbg_J = (float *)malloc(sizeof(float) * 12 * (size_t)m); memcpy(w->dbg_J, J, sizeof(float) * 12 * (size_t)m);
            w->dbg_c = (float *)malloc(sizeof(float) * (size_t)m); memcpy(w->dbg_c, c, sizeof(float) * (size_t)m);
            w->dbg_cfm = (float *)malloc(sizeof(float) * (size_t)m); memcpy(w->dbg_cfm, cfm, sizeof(float) * (size_t)m);
            w->dbg_lo = (float *)malloc(sizeof(float) * (size_t)m); memcpy(w->dbg_lo, lo, sizeof(float) * (size_t)m);
            w->dbg_hi = (float *)malloc(sizeof(float) * (size_t)m); memcpy(w->dbg_hi, hi, sizeof(float) * (size_t)m);
            w->dbg_rhs = (float *)malloc(sizeof(float) * (size_t)m); memcpy(w->dbg_rhs, rhs, sizeof(float) * (size_t)m);
            w->dbg_jb = (int *)malloc(sizeof(int) * 2 * (size_t)m); memcpy(w->dbg_jb, jb, sizeof(int) * 2 * (size_t)m);
        }
        float *fc = (float *)calloc(6 * (size_t)nb, sizeof(float));
        if (order_mode == 3) {
            /* dWorldStep's answer: the exact solution of  A lambda = rhs + w,  lo <= lambda <= hi,
             * A = J invM J^T + cfm/h  (what libode's Dantzig solver dSolveLCP returns; A is SPD for
             * cfm > 0 so the solution is unique and any exact method finds it).  Double precision. */
            double *A = (double *)calloc((size_t)m * (size_t)m, sizeof(double));
            double *bb = (double *)malloc(sizeof(double) * (size_t)m);
            double *x = (double *)calloc((size_t)m, sizeof(double));
            /* M^-1 J^T in double from the float rows and the float inverse masses / world inverse inertias: with the
             * float iMJ of the sweeps, A carries 1e-7 relative errors, which a resting box's nearly singular contact
             * block (four coplanar contacts, cfm/h on the diagonal) amplifies to 1e-3 m/s in the answer -- found by
             * fuzzing the engine's exact dWorldStep against this mode */
            double *iMJd = (double *)calloc((size_t)m * 12, sizeof(double));
            for (int i = 0; i < m; i++) {
                double *im = iMJd + 12 * i;
                const float *Ji = J + 12 * i;
                for (int s1 = 0; s1 < 2; s1++) {
                    int bi = jb[2 * i + s1];
                    if (bi < 0) continue;
                    const float *iI = invI + 12 * bi;
                    for (int k = 0; k < 3; k++) im[6 * s1 + k] = (double)w->b[bi].invMass * (double)Ji[6 * s1 + k];
                    for (int r = 0; r < 3; r++)
                        im[6 * s1 + 3 + r] = (double)iI[4 * r] * (double)Ji[6 * s1 + 3] + (double)iI[4 * r + 1] * (double)Ji[6 * s1 + 4] +
                                             (double)iI[4 * r + 2] * (double)Ji[6 * s1 + 5];
                }
            }
            for (int i = 0; i < m; i++) {
                const double *im = iMJd + 12 * i;
                for (int j2 = 0; j2 < m; j2++) {
                    const float *Jj = J + 12 * j2;
                    double a = 0;
                    for (int s1 = 0; s1 < 2; s1++) {
                        int bi = jb[2 * i + s1];
                        if (bi < 0) continue;
                        for (int s2 = 0; s2 < 2; s2++)
                            if (jb[2 * j2 + s2] == bi)
                                for (int k = 0; k < 6; k++) a += im[6 * s1 + k] * (double)Jj[6 * s2 + k];
                    }
                    A[(size_t)i * m + j2] = a;
                }
                A[(size_t)i * m + i] += (double)cfm[i];
                bb[i] = (double)rhs[i];
                if (findex[i] >= 0) { fprintf(stderr, "ode_oracle: exact mode does not take findex rows\n"); abort(); }
            }
            if (!solve_blcp(m, A, bb, lo, hi, x)) { fprintf(stderr, "ode_oracle: exact LCP did not terminate\n"); abort(); }
            double *fcd = (double *)calloc(6 * (size_t)nb, sizeof(double));
            for (int i = 0; i < m; i++) {
                lambda[i] = (float)x[i];
                const double *im = iMJd + 12 * i;
                int b1 = jb[2 * i], b2 = jb[2 * i + 1];
                for (int k = 0; k < 6; k++) fcd[6 * b1 + k] += x[i] * im[k];
                if (b2 >= 0) for (int k = 0; k < 6; k++) fcd[6 * b2 + k] += x[i] * im[6 + k];
            }
            for (size_t k = 0; k < 6 * (size_t)nb; k++) fc[k] = (float)fcd[k];
            free(A); free(bb); free(x); free(iMJd); free(fcd);
        }
        /* SOR_LCP */
        for (int i = 0; i < m && order_mode != 3; i++) {
            int b2 = jb[2 * i + 1];
            float *im = iMJ + 12 * i, *Ji = J + 12 * i;
            float sum = 0;
            for (int k = 0; k < 6; k++) sum += im[k] * Ji[k];
            if (b2 >= 0) for (int k = 0; k < 6; k++) sum += im[6 + k] * Ji[6 + k];
            Ad[i] = w->sor_w / (sum + cfm[i]);
            for (int k = 0; k < 12; k++) Ji[k] *= Ad[i];
            rhs[i] *= Ad[i];
            Adcfm[i] = Ad[i] * cfm[i];
        }
        int *order = (int *)malloc(sizeof(int) * (size_t)m);
        if (order_mode == 2 && perm) memcpy(order, perm, sizeof(int) * (size_t)m);
        else if (order_mode == 1) { for (int i = 0; i < m; i++) order[i] = i; }
        else {
            int head = 0, tail = m - 1;
            for (int i = 0; i < m; i++) { if (findex[i] < 0) order[head++] = i; else order[tail--] = i; }
        }
        for (int it = 0; it < w->iters && order_mode != 3; it++) {
            if (order_mode == 0 && (it & 7) == 0) {
                for (int i = 1; i < m; i++) {
                    int t = order[i], sw = ode_rand_int(w, i + 1);
                    order[i] = order[sw]; order[sw] = t;
                }
            }
            for (int ii = 0; ii < m; ii++) {
                int i = order[ii];
                int b1 = jb[2 * i], b2 = jb[2 * i + 1];
                float old_lambda = lambda[i];
                float delta = rhs[i] - old_lambda * Adcfm[i];
                float *fc1 = fc + 6 * b1, *Ji = J + 12 * i;
                float *fc2 = 0;
                if (w->perturb_fma) {
                    /* sensitivity experiment only (tests/test_oracle_kat.py): contract the dot products with
                     * fused multiply-adds, the rounding a GPU build with FMA contraction would produce */
                    float a1 = fc1[0] * Ji[0];
                    for (int k = 1; k < 6; k++) a1 = fmaf(fc1[k], Ji[k], a1);
                    delta -= a1;
                    if (b2 >= 0) {
                        fc2 = fc + 6 * b2;
                        float a2 = fc2[0] * Ji[6];
                        for (int k = 1; k < 6; k++) a2 = fmaf(fc2[k], Ji[6 + k], a2);
                        delta -= a2;
                    }
                } else {
                delta -= fc1[0] * Ji[0] + fc1[1] * Ji[1] + fc1[2] * Ji[2] + fc1[3] * Ji[3] + fc1[4] * Ji[4] + fc1[5] * Ji[5];
                if (b2 >= 0) {
                    fc2 = fc + 6 * b2;
                    delta -= fc2[0] * Ji[6] + fc2[1] * Ji[7] + fc2[2] * Ji[8] + fc2[3] * Ji[9] + fc2[4] * Ji[10] + fc2[5] * Ji[11];
                }
                }
                float hi_act, lo_act;
                if (findex[i] >= 0) { hi_act = fabsf(hi[i] * lambda[findex[i]]); lo_act = -hi_act; }
                else { hi_act = hi[i]; lo_act = lo[i]; }
                float new_lambda = old_lambda + delta;
                if (new_lambda < lo_act) { delta = lo_act - old_lambda; lambda[i] = lo_act; }
                else if (new_lambda > hi_act) { delta = hi_act - old_lambda; lambda[i] = hi_act; }
                else lambda[i] = new_lambda;
                const float *im = iMJ + 12 * i;
                if (w->perturb_fma) {
                    for (int k = 0; k < 6; k++) fc1[k] = fmaf(delta, im[k], fc1[k]);
                    if (fc2) for (int k = 0; k < 6; k++) fc2[k] = fmaf(delta, im[6 + k], fc2[k]);
                } else {
                for (int k = 0; k < 6; k++) fc1[k] += delta * im[k];
                if (fc2) for (int k = 0; k < 6; k++) fc2[k] += delta * im[6 + k];
                }
            }
        }
        /* velocity update: v += h * cforce */
        for (int i = 0; i < nb; i++) {
            obody *b = &w->b[i];
            for (int k = 0; k < 3; k++) b->lvel[k] += h * fc[6 * i + k];
            for (int k = 0; k < 3; k++) b->avel[k] += h * fc[6 * i + 3 + k];
        }
        free(fc); free(order); free(J); free(iMJ); free(c); free(cfm); free(lo); free(hi);
        free(rhs); free(Ad); free(Adcfm); free(findex); free(jb);
    }
    free(jrow0);

    /* v += h * invM * fe, then dxStepBody; zero the accumulators */
    for (int i = 0; i < nb; i++) {
        obody *b = &w->b[i];
        for (int k = 0; k < 3; k++) b->lvel[k] += h * b->invMass * b->facc[k];
        float th[3] = {b->tacc[0] * h, b->tacc[1] * h, b->tacc[2] * h}, t[3];
        mul0_331(t, invI + 12 * i, th);
        for (int k = 0; k < 3; k++) b->avel[k] += t[k];
        for (int k = 0; k < 3; k++) b->pos[k] += h * b->lvel[k];
        /* dWtoDQ + infinitesimal rotation */
        const float *wv = b->avel, *q = b->q;
        float dq[4];
        dq[0] = 0.5f * (-wv[0] * q[1] - wv[1] * q[2] - wv[2] * q[3]);
        dq[1] = 0.5f * (wv[0] * q[0] + wv[1] * q[3] - wv[2] * q[2]);
        dq[2] = 0.5f * (-wv[0] * q[3] + wv[1] * q[0] + wv[2] * q[1]);
        dq[3] = 0.5f * (wv[0] * q[2] - wv[1] * q[1] + wv[2] * q[0]);
        for (int k = 0; k < 4; k++) b->q[k] += h * dq[k];
        normalize4(b->q);
        orc_q_to_r(b->q, b->R);
        memset(b->facc, 0, 12); memset(b->tacc, 0, 12);
    }
    free(invI);
    return 1;
}

/* ------------------------------------------------------------------ snapshot pack */

/* GetTransformMat, /root/reference/src/main.c:602-622: column-major 4x4 = transpose of ODE's R */
static void pack_transform(const float *pos, const float *rot, float res[16]) {
    res[0] = rot[0]; res[1] = rot[4]; res[2] = rot[8]; res[3] = 0;
    res[4] = rot[1]; res[5] = rot[5]; res[6] = rot[9]; res[7] = 0;
    res[8] = rot[2]; res[9] = rot[6]; res[10] = rot[10]; res[11] = 0;
    res[12] = pos[0]; res[13] = pos[1]; res[14] = pos[2]; res[15] = 1;
}
void orc_pack_body_transform(const orc_world *w, int b, float out16[16]) { pack_transform(w->b[b].pos, w->b[b].R, out16); }
void orc_pack_geom_transform(const orc_world *w, int g, float out16[16]) {
    const ogeom *ge = &w->g[g];
    if (ge->body >= 0) pack_transform(w->b[ge->body].pos, w->b[ge->body].R, out16);
    else pack_transform(ge->pos, ge->R, out16);
}
