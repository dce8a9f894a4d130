/* physics_server.c -- headless host side of the reference's physics server, in plain C.
 *
 * Mirrors, call for call, what StartServer() does with ODE in /root/reference/src/main.c
 * (world setup :94-98, static map :115-121 via AddBodyMap :735-761, spawned bodies via AddBody
 * :695-733, the fixed-step tick :206-216 with NearCallback :674-693, and the snapshot pack
 * :218-243 with GetTransformMat :602-622) -- without enet, raylib or the GUI.  It compiles against
 * this repo's <ode/ode.h> and links libode_b200.so, i.e. it is the drop-in test of the boundary:
 * the only thing that changed for the host code is the library behind the ODE names.
 *
 * usage: physics_server <seed> <n_spawn> <n_kinematic> <ticks> <dt> <mode: compat|device|wire> <out.bin> [step|quick]
 * (step, the default: dWorldStep as the reference calls it -- the library solves the step's LCP exactly; quick: the
 * same call made to run dWorldQuickStep's 20 sweeps, dWorldSetStepSolverB200(world, -1, 0))
 * (wire = device-resident tick + the MsgUpdateBodies image packed on the GPU, dWorldPackMsgUpdateBodiesB200)
 * Writes the MsgUpdateBodies image (inc/msgs.h:30-33) after the last tick to out.bin.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ode/ode.h"
#include "ode_b200.h"

#define MAX_BODIES 512 /* inc/body.h:6 */

typedef enum { CMASK_MAP = 1, CMASK_OBJ = 2, CMASK_ALL = ~0 } CollMask;      /* inc/body.h:8-12 */
typedef enum { BODYTYPE_NULL, BODYTYPE_SPHERE, BODYTYPE_BOX } BodyType;      /* inc/body.h:14-18 */
typedef enum {                                                                /* inc/msgs.h:6-13 */
    MSGTYPE_C_PLAYER_ID, MSGTYPE_C_UPDATE_PLAYERS, MSGTYPE_S_PLAYER_UPDATE, MSGTYPE_C_UPDATE_BODIES, MSGTYPE_S_NEW_BODY
} MsgType;

typedef struct { float x, y, z; } Vec3;                 /* raylib Vector3 */
typedef struct { unsigned char r, g, b, a; } Color4;    /* raylib Color */

typedef struct { dBodyID body; dGeomID geom; BodyType type; } Body;                       /* inc/body.h:20-24 */
typedef struct { BodyType type; dReal transform[16]; Vec3 size; Color4 col; } BodyState;   /* inc/body.h:26-31 */
typedef struct { MsgType msg; BodyState bodies[MAX_BODIES]; } MsgUpdateBodies;             /* inc/msgs.h:30-33 */

static dWorldID world;
static dSpaceID space;
static dJointGroupID contactGroup;

/* PRNG of src/rand.c:7-34 */
static uint32_t randState = 0;
static uint32_t Rand_Next(void) {
    randState += 0xE120FC15u;
    uint64_t t = (uint64_t)randState * 0x4A39B70Du;
    const uint32_t m1 = (uint32_t)((t >> 32) ^ t);
    t = (uint64_t)m1 * 0x12FAD5C9u;
    return (uint32_t)((t >> 32) ^ t);
}
static int Rand_Int(int lo, int hi) { return (int)(Rand_Next() % (uint32_t)(hi - lo)) + lo; }
static double Rand_Double(double lo, double hi) { return lo + Rand_Next() / (double)0xFFFFFFFFu * (hi - lo); }
static Color4 Rand_Color(unsigned char lo, unsigned char hi) {
    Color4 c;
    c.r = (unsigned char)Rand_Int(lo, hi); c.g = (unsigned char)Rand_Int(lo, hi); c.b = (unsigned char)Rand_Int(lo, hi);
    c.a = 255;
    return c;
}

/* src/main.c:602-622 */
static void GetTransformMat(dReal res[16], const dReal *pos, const dReal *rot) {
    for (int c = 0; c < 3; c++) {
        res[4 * c + 0] = rot[c];
        res[4 * c + 1] = rot[4 + c];
        res[4 * c + 2] = rot[8 + c];
        res[4 * c + 3] = 0.0f;
    }
    res[12] = pos[0]; res[13] = pos[1]; res[14] = pos[2]; res[15] = 1.0f;
}

/* src/main.c:624-651 (with its :639 quirk) */
static void GetTransformMatV(dReal res[16], Vec3 pos, Vec3 rot) {
    const dReal cx = cos(rot.x), sx = sin(rot.x), cy = cos(rot.y), sy = sin(rot.y), cz = cos(rot.z), sz = sin(rot.z);
    res[0] = cy * cz; res[1] = cz * sx * sy - cx * sz; res[2] = cx * cz * sy + sx * sz; res[3] = 0.0f;
    res[4] = cy * sz; res[5] = cx * cz + sx * sy * sz; res[6] = -cz * sx + cx * sy * sx; res[7] = 0.0f;
    res[8] = -sy; res[9] = cy * sx; res[10] = cx * cy; res[11] = 0.0f;
    res[12] = pos.x; res[13] = pos.y; res[14] = pos.z; res[15] = 1.0f;
}

/* src/main.c:674-693 */
static void NearCallback(void *data, dGeomID o1, dGeomID o2) {
    (void)data;
    enum { MAX_CONTACTS = 8 };
    dContact contacts[MAX_CONTACTS];
    const int nc = dCollide(o1, o2, MAX_CONTACTS, &contacts[0].geom, sizeof(dContact));
    if (nc <= 0) return;
    for (int i = 0; i < nc; i++) {
        contacts[i].surface.mode = dContactBounce;
        contacts[i].surface.bounce = 0.2f;
        contacts[i].surface.bounce_vel = 0.1f;
        contacts[i].surface.mu = dInfinity;
        dJointID c = dJointCreateContact(world, contactGroup, &contacts[i]);
        dJointAttach(c, dGeomGetBody(o1), dGeomGetBody(o2));
    }
}

/* src/main.c:695-733 */
static int AddBody(Body *bodies, BodyState *states, CollMask category, CollMask collide, BodyState state, char isKinematic) {
    for (int i = 0; i < MAX_BODIES; i++) {
        if (bodies[i].type != BODYTYPE_NULL) continue;
        Body *body = &bodies[i];
        body->type = state.type;
        body->body = dBodyCreate(world);
        dReal pos[3], rm[12];
        for (int k = 0; k < 3; k++) pos[k] = state.transform[12 + k];
        for (int k = 0; k < 12; k++) rm[k] = state.transform[k];
        dBodySetPosition(body->body, pos[0], pos[1], pos[2]);
        dBodySetRotation(body->body, rm);
        if (isKinematic) dBodySetKinematic(body->body);
        switch (state.type) {
            case BODYTYPE_SPHERE: body->geom = dCreateSphere(space, state.size.x); break;
            case BODYTYPE_BOX: body->geom = dCreateBox(space, state.size.x, state.size.y, state.size.z); break;
            default: return -1;
        }
        dGeomSetCategoryBits(body->geom, (unsigned long)category);
        dGeomSetCollideBits(body->geom, (unsigned long)collide);
        dGeomSetBody(body->geom, body->body);
        states[i] = state;
        return i;
    }
    return -1;
}

/* src/main.c:735-761 (including the doubled dGeomSetCategoryBits of :751-752) */
static int AddBodyMap(Body *bodies, BodyState *states, Vec3 pos, Vec3 rot, Vec3 size, Color4 col) {
    for (int i = 0; i < MAX_BODIES; i++) {
        if (bodies[i].type != BODYTYPE_NULL) continue;
        Body *body = &bodies[i];
        body->type = BODYTYPE_BOX;
        body->geom = dCreateBox(space, size.x, size.y, size.z);
        dReal trans[16], rm[12];
        GetTransformMatV(trans, pos, rot);
        for (int k = 0; k < 12; k++) rm[k] = trans[k];
        dGeomSetPosition(body->geom, pos.x, pos.y, pos.z);
        dGeomSetRotation(body->geom, rm);
        dGeomSetCategoryBits(body->geom, (unsigned long)(uint32_t)CMASK_MAP);
        dGeomSetCategoryBits(body->geom, (unsigned long)(uint32_t)(CMASK_ALL & ~CMASK_MAP));
        body->body = NULL;
        memset(&states[i], 0, sizeof(states[i]));
        states[i].size = size; states[i].col = col; states[i].type = BODYTYPE_BOX;
        memcpy(states[i].transform, trans, sizeof(dReal) * 16);
        return i;
    }
    return -1;
}

int main(int argc, char **argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: %s seed n_spawn n_kinematic ticks dt compat|device|wire out.bin [step|quick] [warm-up ticks]\n", argv[0]);
        return 2;
    }
    randState = (uint32_t)strtoul(argv[1], 0, 10);
    const int n_spawn = atoi(argv[2]), n_kin = atoi(argv[3]), ticks = atoi(argv[4]);
    const float dt = (float)atof(argv[5]);
    const int wire_mode = strcmp(argv[6], "wire") == 0;
    const int device_mode = wire_mode || strcmp(argv[6], "device") == 0;

    dInitODE();                                  /* src/main.c:94-98 */
    world = dWorldCreate();
    dWorldSetGravity(world, 0.0, -9.8, 0.0);
    space = dHashSpaceCreate(0);
    contactGroup = dJointGroupCreate(0);
    if (argc > 8 && strcmp(argv[8], "quick") == 0) dWorldSetStepSolverB200(world, -1, 0.f); /* dWorldStep := dWorldQuickStep */
    const int warm = argc > 9 ? atoi(argv[9]) : 0; /* untimed ticks before the timed ones (bench.py --workload C1) */

    static Body bodies[MAX_BODIES];
    static BodyState bodyStates[MAX_BODIES];
    for (int i = 0; i < MAX_BODIES; i++) bodies[i].type = bodyStates[i].type = BODYTYPE_NULL;

    const Color4 grey = {80, 80, 80, 255};
    AddBodyMap(bodies, bodyStates, (Vec3){0.f, 0.f, 0.f}, (Vec3){0.f, 0.f, 0.f}, (Vec3){100.f, 1.f, 100.f}, grey); /* :115 */
    AddBodyMap(bodies, bodyStates, (Vec3){4.f, 3.f, 0.f}, (Vec3){0.f, 0.f, -0.5f}, (Vec3){0.5f, 8.f, 12.f}, grey); /* :118 */
    AddBodyMap(bodies, bodyStates, (Vec3){0.f, 3.f, 6.f}, (Vec3){0.f, 0.f, 0.f}, (Vec3){12.f, 8.f, 0.5f}, grey);   /* :120 */
    AddBodyMap(bodies, bodyStates, (Vec3){0.f, 3.f, -6.f}, (Vec3){0.f, 0.f, 0.f}, (Vec3){12.f, 8.f, 0.5f}, grey);  /* :121 */

    /* the client's `M` key, src/main.c:502-522, n_spawn times; server side :178-182 */
    for (int s = 0; s < n_spawn; s++) {
        Vec3 pos;
        pos.x = (float)Rand_Double(-4.0, 4.0);
        pos.y = (float)Rand_Double(20.0, 50.0);
        pos.z = (float)Rand_Double(-4.0, 4.0);
        BodyState state;
        memset(&state, 0, sizeof(state));
        if (Rand_Int(0, 2) == 0) {
            state.type = BODYTYPE_BOX;
            state.size.x = (float)Rand_Double(0.2, 1.0);
            state.size.y = (float)Rand_Double(0.2, 1.0);
            state.size.z = (float)Rand_Double(0.2, 1.0);
        } else {
            state.type = BODYTYPE_SPHERE;
            state.size.x = (float)Rand_Double(0.1, 0.4);
        }
        state.col = Rand_Color(30, 190);
        GetTransformMatV(state.transform, pos, (Vec3){0.f, 0.f, 0.f});
        AddBody(bodies, bodyStates, CMASK_OBJ, CMASK_OBJ | CMASK_MAP, state, 0);
    }
    /* kinematic "player" spheres (BASELINE.json config 1; radius from src/main.c:315) */
    for (int p = 0; p < n_kin; p++) {
        BodyState state;
        memset(&state, 0, sizeof(state));
        state.type = BODYTYPE_SPHERE;
        state.size.x = 0.5f;
        GetTransformMatV(state.transform, (Vec3){0.f + 1.5f * p, 2.f, -3.f}, (Vec3){0.f, 0.f, 0.f});
        AddBody(bodies, bodyStates, CMASK_OBJ, CMASK_OBJ | CMASK_MAP, state, 1);
    }

    struct timespec t0, t1;
    for (int t = 0; t < warm + ticks; t++) {     /* src/main.c:211-216 */
        if (t == warm) { dWorldWaitB200(world); clock_gettime(CLOCK_MONOTONIC, &t0); }
        if (device_mode) dSpaceCollideDeviceB200(space, 8);
        else dSpaceCollide(space, NULL, NearCallback);
        dWorldStep(world, dt);
        dJointGroupEmpty(contactGroup);
    }
    dWorldWaitB200(world);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms_per_tick = ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6) / (ticks > 0 ? ticks : 1);
    dStepStatsB200 stats;
    dWorldGetStatsB200(world, &stats);

    static MsgUpdateBodies updatedBodies;        /* src/main.c:239-240 */
    if (wire_mode) {
        /* the whole of src/main.c:221-240 as one kernel + one copy */
        static dBodyID slot_body[MAX_BODIES];
        static dGeomID slot_geom[MAX_BODIES];
        static int slot_type[MAX_BODIES];
        static float slot_size[MAX_BODIES][3];
        static unsigned slot_col[MAX_BODIES];
        for (int i = 0; i < MAX_BODIES; i++) {
            slot_type[i] = bodyStates[i].type;
            slot_body[i] = bodies[i].type != BODYTYPE_NULL ? bodies[i].body : NULL;
            slot_geom[i] = bodies[i].type != BODYTYPE_NULL ? bodies[i].geom : NULL;
            slot_size[i][0] = bodyStates[i].size.x; slot_size[i][1] = bodyStates[i].size.y; slot_size[i][2] = bodyStates[i].size.z;
            memcpy(&slot_col[i], &bodyStates[i].col, 4);
        }
        dWorldBindSnapshotSlotsB200(world, MAX_BODIES, slot_body, slot_geom, slot_type, &slot_size[0][0], slot_col);
        const size_t nbytes = dWorldPackMsgUpdateBodiesB200(world, &updatedBodies, MSGTYPE_C_UPDATE_BODIES, 1);
        if (nbytes != sizeof(updatedBodies)) { fprintf(stderr, "wire image size %zu != %zu\n", nbytes, sizeof(updatedBodies)); return 1; }
    } else {
    for (int i = 0; i < MAX_BODIES; i++) {       /* src/main.c:221-237 */
        if (BODYTYPE_NULL == bodies[i].type) continue;
        const dReal *pos, *rot;
        if (bodies[i].body) {
            pos = dBodyGetPosition(bodies[i].body);
            rot = dBodyGetRotation(bodies[i].body);
        } else {
            pos = dGeomGetPosition(bodies[i].geom);
            rot = dGeomGetRotation(bodies[i].geom);
        }
        GetTransformMat(bodyStates[i].transform, pos, rot);
    }
    memset(&updatedBodies, 0, sizeof(updatedBodies));
    updatedBodies.msg = MSGTYPE_C_UPDATE_BODIES;
    memcpy(updatedBodies.bodies, bodyStates, sizeof(bodyStates));
    }
    FILE *f = fopen(argv[7], "wb");
    if (!f) { perror("out"); return 1; }
    fwrite(&updatedBodies, sizeof(updatedBodies), 1, f);
    fclose(f);
    printf("ticks=%d bytes=%zu ms_per_tick=%.6f exact_status=%d islands=%d max_island_rows=%d contacts=%d\n", ticks,
           sizeof(updatedBodies), ms_per_tick, stats.exact_status, stats.n_islands, stats.max_island_rows, stats.n_contacts);

    for (int i = 0; i < MAX_BODIES; i++) {       /* src/main.c:259-267 */
        if (bodies[i].type == BODYTYPE_NULL) continue;
        if (bodies[i].body) dBodyDestroy(bodies[i].body);
        dGeomDestroy(bodies[i].geom);
    }
    dJointGroupDestroy(contactGroup);
    dWorldDestroy(world);
    dCloseODE();
    return 0;
}
