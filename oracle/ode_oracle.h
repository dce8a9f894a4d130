/* ode_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU (plain C99, float32, -ffp-contract=off) restatement of the slice of libode that the
 * reference reaches from src/main.c:94-98,206-243,674-761: hash-space broadphase, dCollide for
 * sphere/box/plane/trimesh pairs, contact-joint rows, QuickStep SOR-PGS, dxStepBody and the
 * reference's own GetTransformMat snapshot pack.
 *
 * PARITY UNPINNED: libode is an un-vendored, un-versioned dependency of the reference
 * (`#include "ode/ode.h"`, src/main.c:11) and is absent from /root/reference and from this
 * image, and the reference has no tests or golden vectors.  The restatement follows the
 * published ODE 0.13-0.16 algorithm (dSINGLE) as recorded in SURVEY.md Appendix A and is pinned
 * only by hand-derivable known-answer cases (tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this library.  The product (libode_b200.so) never links or loads it.
 */
#ifndef ODE_ORACLE_H
#define ODE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SPHERE = 0, ORC_BOX = 1, ORC_PLANE = 4, ORC_TRIMESH = 8 };

enum { ORC_BODY_KINEMATIC = 1, ORC_BODY_NOGRAVITY = 2, ORC_BODY_GYRO = 4 };

/* surface mode bits = ODE's dContact* values */
enum {
    ORC_MU2 = 0x001, ORC_FDIR1 = 0x002, ORC_BOUNCE = 0x004, ORC_SOFT_ERP = 0x008,
    ORC_SOFT_CFM = 0x010, ORC_MOTION1 = 0x020, ORC_MOTION2 = 0x040, ORC_MOTIONN = 0x080,
    ORC_SLIP1 = 0x100, ORC_SLIP2 = 0x200, ORC_APPROX1_1 = 0x1000, ORC_APPROX1_2 = 0x2000
};

typedef struct orc_contact_geom {
    float pos[3];
    float normal[3]; /* from g2 into g1 */
    float depth;
    int g1, g2;
    int side1, side2;
} orc_contact_geom;

typedef struct orc_surface {
    int mode;
    float mu, mu2, bounce, bounce_vel, soft_erp, soft_cfm;
    float motion1, motion2, motionN, slip1, slip2;
    float fdir1[3];
} orc_surface;

typedef struct orc_world orc_world;

orc_world *orc_create(void);
void orc_destroy(orc_world *);
void orc_set_gravity(orc_world *, float x, float y, float z);
void orc_set_params(orc_world *, float erp, float cfm, int iters, float sor_w);
void orc_set_contact_params(orc_world *, float max_vel, float min_depth);
/* sensitivity experiment: solve with FMA-contracted dot products instead of ODE's rounding */
void orc_set_perturb_fma(orc_world *, int on);

/* bodies: default mass 1, I = identity (dBodyCreate). q = (w,x,y,z). R optional (NULL -> dQtoR(q)).
 * inertia: 9 floats row-major (NULL -> identity). returns body index */
int orc_add_body(orc_world *, const float pos[3], const float q[4], const float *R12,
                 const float lvel[3], const float avel[3], float mass, const float *inertia9,
                 int flags, int env);
int orc_num_bodies(const orc_world *);
void orc_get_body(const orc_world *, int b, float pos[3], float q[4], float R12[12],
                  float lvel[3], float avel[3]);
void orc_set_body_state(orc_world *, int b, const float pos[3], const float q[4],
                        const float *R12, const float lvel[3], const float avel[3]);
void orc_add_force(orc_world *, int b, const float f[3], const float t[3]);

/* geoms: dims = sphere (r), box (lx,ly,lz full lengths), plane (a,b,c,d), trimesh (mesh id).
 * body = -1 for static geoms, which then use pos / R12 (row-major 3x4). env = -1: all envs */
int orc_add_mesh(orc_world *, const float *verts, int nverts, const int *tris, int ntris);
int orc_add_geom(orc_world *, int type, const float dims[4], int body, const float pos[3],
                 const float *R12, unsigned cat, unsigned col, int env);
int orc_num_geoms(const orc_world *);
void orc_get_aabb(orc_world *, int g, float aabb[6]);

/* broadphase: every pair passing ODE's collideAABBs filter, as (min id, max id), sorted.
 * method 0 = multi-resolution hash space (restating dxHashSpace::collide), 1 = brute force.
 * returns the number of pairs (which may exceed cap; only cap are written). */
long orc_broadphase(orc_world *, int method, int *pairs, long cap);

/* dCollide(g1, g2, maxc): contacts in ODE's order; returns count */
int orc_collide(orc_world *, int g1, int g2, int maxc, orc_contact_geom *out);

/* dJointCreateContact + dJointAttach(b1, b2) (either may be -1 = NULL body) */
int orc_add_contact(orc_world *, const orc_contact_geom *, const orc_surface *, int b1, int b2);
void orc_clear_contacts(orc_world *);
int orc_num_contacts(const orc_world *);

/* reference's tick (src/main.c:212-214 with NearCallback src/main.c:674-693): collide every
 * broadphase pair with maxc contacts, create a contact joint per contact with `surf`. Returns
 * number of contacts created. Pairs are visited in sorted order. */
long orc_collide_all(orc_world *, int maxc, const orc_surface *surf);

/* dWorldQuickStep. order_mode 0: ODE ordering (findex<0 first, reshuffle every 8 iterations with
 * ODE's LCG); 1: fixed order = rows in joint order every iteration (joint j -> rows 3j..);
 * 2: caller permutation `perm` of row indices (length = number of rows);
 * 3: no sweeps -- the exact solution of the step's LCP (what dWorldStep's Dantzig solver returns),
 *    by principal pivoting in double precision; rows with findex (Approx1 friction) are not accepted. */
int orc_quickstep(orc_world *, float h, int order_mode, const int *perm);
int orc_num_rows(const orc_world *);
/* diagnostics of the last step: max |delta lambda| of last iteration etc. */
void orc_last_lambda(const orc_world *, float *lambda, int n);
/* test diagnostics: keep the rows of each step (J 12 floats, c, cfm/h, lo, hi, rhs, body pair per row,
 * before SOR_LCP scales them by Ad) so that tests can evaluate constraint residuals and re-solve the
 * same LCP with an independently written solver.  Returns the number of rows of the last step. */
void orc_set_keep_rows(orc_world *, int on);
int orc_last_rows(const orc_world *, float *J, float *c, float *cfm, float *lo, float *hi, float *rhs, int *jb, int n);

/* reference's snapshot pack: GetTransformMat (src/main.c:602-622) per body/geom */
void orc_pack_body_transform(const orc_world *, int b, float out16[16]);
void orc_pack_geom_transform(const orc_world *, int g, float out16[16]);

/* helpers exposed for unit tests */
void orc_q_to_r(const float q[4], float R12[12]);
void orc_r_to_q(const float R12[12], float q[4]);
void orc_plane_space(const float n[3], float p[3], float q[3]);
unsigned orc_rand_next(unsigned *state); /* reference PRNG, src/rand.c:7-13 */

#ifdef __cplusplus
}
#endif
#endif
