/* slab_server.c -- a C host for the slab-decomposed world (BASELINE config 5).
 *
 * The reference's server is a C program that owns one ODE world (/root/reference/src/main.c:59-270).  This is the same
 * kind of program for a world too large for one GPU: one process per GPU, each building ITS slab of a lattice pile
 * through the ODE handle API (dBodyCreate / dCreateBox / dGeomSetBody, as AddBody does, src/main.c:695-733), then
 * handing the world to the slab driver inside libode_b200.so (dSlabCreateB200): the per-tick halo exchange runs over
 * NCCL inside the library, ordered by CUDA events; the host calls dSlabTickB200 and, every 16 ticks,
 * dSlabMigrateB200.  The only thing the application supplies besides the world is the 128-byte NCCL unique id, which
 * rank 0 obtains from dSlabGetUniqueIdB200 and hands to the others -- here through a file.
 *
 * usage:
 *   one process per GPU:  slab_server nccl  <rank> <n_ranks> <id-file> <cols> <nz> <ny> <ticks>
 *   one process, one GPU: slab_server local <n_slabs>               <cols> <nz> <ny> <ticks>
 * (local: the slabs of ONE process are connected with dSlabConnectLocalB200 and ticked by dSlabTickLocalB200 -- same
 * phases, device copies as transport; this is what the single-GPU tests run)
 * Prints one line per slab: rank, bodies owned, migrated in / out, halo bodies selected, overflows, ms per tick, and
 * the owned bodies' momentum, lowest point and top speed (a pile at rest on the plane: the sanity check of the run).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "ode/ode.h"
#include "ode_b200.h"

enum { CAT_MAP = 1, CAT_OBJ = 2, CAT_GHOST = 4 };
#define SPACING 1.8f
#define MARGIN_COLS 4
#define MIG_CAP 4096

typedef struct {
    dWorldID world;
    dSpaceID space;
    dSlabID slab;
    int rank, n_own, n_bodies, pool;
} Slab;

static uint32_t rng_state;
static uint32_t rng_next(void) { /* src/rand.c:7-13 */
    rng_state += 0xE120FC15u;
    uint64_t t = (uint64_t)rng_state * 0x4A39B70Du;
    const uint32_t m1 = (uint32_t)((t >> 32) ^ t);
    t = (uint64_t)m1 * 0x12FAD5C9u;
    return (uint32_t)((t >> 32) ^ t);
}
static float rng_float(float lo, float hi) { return lo + (float)(rng_next() / (double)0xFFFFFFFFu) * (hi - lo); }

/* the slab `rank` of `n_ranks`: columns [rank * cols, (rank + 1) * cols) of a (n_ranks * cols) x nz x ny lattice of boxes
 * and spheres drawn like the reference's spawns (src/main.c:504-521), five planes around it, and -- unless it is the last
 * slab -- a pool of parked kinematic ghost slots for the upper neighbour's boundary bodies */
static void build_slab(Slab *s, int rank, int n_ranks, int cols, int nz, int ny, const char *nccl_id) {
    memset(s, 0, sizeof(*s));
    s->rank = rank;
    s->world = dWorldCreate();
    dWorldSetGravity(s->world, 0, -9.8f, 0);
    dWorldSetQuickStepNumIterations(s->world, 20);
    s->space = dHashSpaceCreate(0);
    dSurfaceParameters surf;
    memset(&surf, 0, sizeof(surf));
    surf.mode = dContactBounce; surf.bounce = 0.2f; surf.bounce_vel = 0.1f; surf.mu = dInfinity; /* src/main.c:684-687 */
    dWorldSetSurfaceB200(s->world, &surf);

    const int nx_total = cols * n_ranks;
    const float x0 = -0.5f * (nx_total - 1) * SPACING, z0 = -0.5f * (nz - 1) * SPACING;
    const float half_x = 0.5f * nx_total * SPACING + 1.0f, half_z = 0.5f * nz * SPACING + 1.0f;
    const float planes[5][4] = {{0, 1, 0, 0}, {1, 0, 0, -half_x}, {-1, 0, 0, -half_x}, {0, 0, 1, -half_z}, {0, 0, -1, -half_z}};
    for (int i = 0; i < 5; i++) { /* static geoms first: the geom of body b is 5 + b */
        dGeomID g = dCreatePlane(s->space, planes[i][0], planes[i][1], planes[i][2], planes[i][3]);
        dGeomSetCategoryBits(g, CAT_MAP);
        dGeomSetCollideBits(g, CAT_OBJ);
    }
    rng_state = 5u + 977u * (uint32_t)rank;
    for (int ix = 0; ix < cols; ix++)
        for (int iz = 0; iz < nz; iz++)
            for (int iy = 0; iy < ny; iy++) {
                dBodyID b = dBodyCreate(s->world);
                dBodySetPosition(b, x0 + (rank * cols + ix) * SPACING + rng_float(-0.03f, 0.03f), 1.0f + iy * SPACING,
                                 z0 + iz * SPACING + rng_float(-0.03f, 0.03f));
                /* the two columns next to an inner face start with 4 m/s towards it: they cross it and change owner */
                if (ix >= cols - 2 && rank < n_ranks - 1) dBodySetLinearVel(b, 4.0f, 0, 0);
                if (ix < 2 && rank > 0) dBodySetLinearVel(b, -4.0f, 0, 0);
                dGeomID g;
                if (rng_next() & 1u) g = dCreateBox(s->space, rng_float(0.2f, 1.0f), rng_float(0.2f, 1.0f), rng_float(0.2f, 1.0f));
                else g = dCreateSphere(s->space, rng_float(0.1f, 0.4f));
                dGeomSetBody(g, b);
                dGeomSetCategoryBits(g, CAT_OBJ);
                dGeomSetCollideBits(g, CAT_OBJ | CAT_MAP | CAT_GHOST);
                s->n_own++;
            }
    s->pool = rank < n_ranks - 1 ? (int)(1.5f * MARGIN_COLS * nz * ny) + 64 : 0;
    for (int i = 0; i < s->pool; i++) { /* ghost slots: parked far below, switched on by the halo */
        dBodyID b = dBodyCreate(s->world);
        dBodySetPosition(b, 0, -1000.0f - i, 0);
        dBodySetKinematic(b);
        dGeomID g = dCreateSphere(s->space, 0.1f);
        dGeomSetBody(g, b);
        dGeomSetCategoryBits(g, CAT_GHOST);
        dGeomSetCollideBits(g, 0);
    }
    s->n_bodies = s->n_own + s->pool;

    dSlabLayoutB200 lay;
    memset(&lay, 0, sizeof(lay));
    lay.face_left = x0 + (rank * cols - 0.5f) * SPACING;
    lay.face_right = lay.face_left + cols * SPACING;
    lay.margin = MARGIN_COLS * SPACING;
    lay.hyst = 0.25f * MARGIN_COLS * SPACING;
    lay.n_own = s->n_own;
    lay.n_static = 5;
    lay.pool = s->pool > 0 ? s->pool : (int)(1.5f * MARGIN_COLS * nz * ny) + 64; /* message size: the same on every rank */
    lay.pool_first_body = s->n_own;
    lay.pool_first_geom = 5 + s->n_own;
    lay.mig_cap = MIG_CAP;
    s->slab = dSlabCreateB200(s->world, s->space, rank, n_ranks, nccl_id, &lay);
}

static void report(Slab *s, int ticks, float ms) {
    dSlabInfoB200 inf;
    dSlabGetInfoB200(s->slab, &inf);
    const int n = dWorldGetNumBodiesB200(s->world);
    float *pos = malloc(sizeof(float) * 3 * n), *lv = malloc(sizeof(float) * 3 * n);
    dWorldGetStateB200(s->world, pos, NULL, lv, NULL, NULL);
    double px = 0, py = 0, pz = 0, vmax = 0, ymin = 1e9;
    int live = 0;
    for (int i = 0; i < n; i++) {
        if (i >= s->n_own && i < s->n_own + s->pool) continue; /* ghost slots */
        if (pos[3 * i + 1] < -500.0f) continue;                 /* parked */
        const double v = sqrt((double)lv[3 * i] * lv[3 * i] + (double)lv[3 * i + 1] * lv[3 * i + 1] + (double)lv[3 * i + 2] * lv[3 * i + 2]);
        px += lv[3 * i]; py += lv[3 * i + 1]; pz += lv[3 * i + 2];
        if (v > vmax) vmax = v;
        if (pos[3 * i + 1] < ymin) ymin = pos[3 * i + 1];
        live++;
    }
    printf("slab %d: owned %d live %d migrated_in %ld migrated_out %ld halo_selected %d halo_overflow %d mig_overflow %d "
           "halo_bytes_per_tick %ld ticks %d ms_per_tick %.4f momentum %.3f %.3f %.3f ymin %.3f vmax %.3f\n",
           s->rank, inf.n_owned, live, inf.migrated_in, inf.migrated_out, inf.halo_selected, inf.halo_overflow, inf.mig_overflow,
           inf.halo_bytes_per_tick, ticks, ms, px, py, pz, ymin, vmax);
    free(pos); free(lv);
}

int main(int argc, char **argv) {
    if (argc < 2) goto usage;
    dInitODE();
    if (!strcmp(argv[1], "nccl") && argc == 9) {
        const int rank = atoi(argv[2]), n_ranks = atoi(argv[3]);
        const char *path = argv[4];
        const int cols = atoi(argv[5]), nz = atoi(argv[6]), ny = atoi(argv[7]), ticks = atoi(argv[8]);
        dSetDeviceB200(rank); /* one process per GPU */
        char id[128];
        if (rank == 0) { /* the launcher's channel for the unique id: a file, written whole and then renamed */
            if (!dSlabGetUniqueIdB200(id)) { fprintf(stderr, "slab_server: NCCL is not available\n"); return 2; }
            char tmp[1024];
            snprintf(tmp, sizeof(tmp), "%s.tmp", path);
            FILE *f = fopen(tmp, "wb");
            if (!f || fwrite(id, 1, 128, f) != 128) { perror("slab_server: id file"); return 2; }
            fclose(f);
            rename(tmp, path);
        } else {
            FILE *f = NULL;
            for (int tries = 0; tries < 6000 && !(f = fopen(path, "rb")); tries++) usleep(10000);
            if (!f || fread(id, 1, 128, f) != 128) { fprintf(stderr, "slab_server: no unique id in %s\n", path); return 2; }
            fclose(f);
        }
        Slab s;
        build_slab(&s, rank, n_ranks, cols, nz, ny, id);
        for (int t = 0; t < 16; t++) dSlabTickB200(s.slab, 1.0f / 60.0f, 8); /* warm-up: first collide uploads the world */
        dWorldWaitB200(s.world);
        dWorldTimerStartB200(s.world);
        for (int t = 0; t < ticks; t++) {
            if (t > 0 && t % 16 == 0) dSlabMigrateB200(s.slab);
            dSlabTickB200(s.slab, 1.0f / 60.0f, 8);
        }
        dWorldTimerStopB200(s.world);
        report(&s, ticks, dWorldTimerElapsedB200(s.world) / ticks);
        dSlabDestroyB200(s.slab);
        dWorldDestroy(s.world);
    } else if (!strcmp(argv[1], "local") && argc == 7) {
        const int n = atoi(argv[2]), cols = atoi(argv[3]), nz = atoi(argv[4]), ny = atoi(argv[5]), ticks = atoi(argv[6]);
        if (n < 1 || n > 16) goto usage;
        Slab s[16];
        dSlabID ids[16];
        for (int r = 0; r < n; r++) { build_slab(&s[r], r, n, cols, nz, ny, NULL); ids[r] = s[r].slab; }
        for (int r = 0; r + 1 < n; r++) dSlabConnectLocalB200(ids[r], ids[r + 1]);
        for (int t = 0; t < 16; t++) dSlabTickLocalB200(ids, n, 1.0f / 60.0f, 8);
        for (int r = 0; r < n; r++) dWorldWaitB200(s[r].world);
        dWorldTimerStartB200(s[0].world);
        for (int t = 0; t < ticks; t++) {
            if (t > 0 && t % 16 == 0) dSlabMigrateLocalB200(ids, n);
            dSlabTickLocalB200(ids, n, 1.0f / 60.0f, 8);
        }
        for (int r = 0; r < n; r++) dWorldWaitB200(s[r].world);
        dWorldTimerStopB200(s[0].world);
        const float ms = dWorldTimerElapsedB200(s[0].world) / ticks;
        for (int r = 0; r < n; r++) report(&s[r], ticks, ms);
        for (int r = 0; r < n; r++) { dSlabDestroyB200(s[r].slab); dWorldDestroy(s[r].world); }
    } else goto usage;
    dCloseODE();
    return 0;
usage:
    fprintf(stderr, "usage: slab_server nccl <rank> <n_ranks> <id-file> <cols> <nz> <ny> <ticks>\n"
                    "       slab_server local <n_slabs> <cols> <nz> <ny> <ticks>\n");
    return 1;
}
