/* ode/ode.h -- the drop-in boundary of libode_b200.so.
 *
 * This header replaces the libode header the reference includes at src/main.c:11
 * (`#include "ode/ode.h"`).  It declares, with C linkage and ODE's own names, argument
 * meaning and struct layouts, exactly the subset of the Open Dynamics Engine API that
 *   (a) the reference's physics server calls (src/main.c, call sites cited per function), and
 *   (b) BASELINE.json's north_star adds (QuickStep, dMass*, planes, trimeshes, quaternion and
 *       velocity accessors).
 * Everything behind these entry points is hand-written CUDA for sm_100a; there is no CPU
 * stepping path.  Device-resident / batched extensions live in <ode_b200.h>.
 */
#ifndef ODE_B200_ODE_H
#define ODE_B200_ODE_H

#include "common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ contact structs */

/* dSurfaceParameters.mode flags (ODE contact.h). */
enum {
    dContactMu2 = 0x001,
    dContactFDir1 = 0x002,
    dContactBounce = 0x004, /* reference: src/main.c:684 */
    dContactSoftERP = 0x008,
    dContactSoftCFM = 0x010,
    dContactMotion1 = 0x020,
    dContactMotion2 = 0x040,
    dContactMotionN = 0x080,
    dContactSlip1 = 0x100,
    dContactSlip2 = 0x200,
    dContactRolling = 0x400, /* accepted, ignored (no rolling friction rows) */
    dContactApprox0 = 0x0000,
    dContactApprox1_1 = 0x1000,
    dContactApprox1_2 = 0x2000,
    dContactApprox1_N = 0x4000,
    dContactApprox1 = 0x7000
};

typedef struct dSurfaceParameters {
    int mode;
    dReal mu;
    dReal mu2;
    dReal rho, rho2, rhoN; /* rolling friction: carried for layout parity, unused */
    dReal bounce;
    dReal bounce_vel;
    dReal soft_erp;
    dReal soft_cfm;
    dReal motion1, motion2, motionN;
    dReal slip1, slip2;
} dSurfaceParameters;

typedef struct dContactGeom {
    dVector3 pos;
    dVector3 normal; /* points from g2 into g1 */
    dReal depth;
    dGeomID g1, g2;
    int side1, side2;
} dContactGeom;

/* reference: `dContact contacts[MAX_CONTACTS]` src/main.c:676 */
typedef struct dContact {
    dSurfaceParameters surface;
    dContactGeom geom;
    dVector3 fdir1;
} dContact;

typedef struct dMass {
    dReal mass;
    dVector3 c;
    dMatrix3 I;
} dMass;

/* geom classes (ODE collision.h) */
enum {
    dSphereClass = 0,
    dBoxClass,
    dCapsuleClass,
    dCylinderClass,
    dPlaneClass,
    dRayClass,
    dConvexClass,
    dGeomTransformClass,
    dTriMeshClass,
    dHeightfieldClass
};

/* ------------------------------------------------------------------ library lifetime */
void dInitODE(void);                 /* src/main.c:94  */
int dInitODE2(unsigned int flags);
void dCloseODE(void);                /* src/main.c:267 */

/* ------------------------------------------------------------------ world */
dWorldID dWorldCreate(void);                                   /* src/main.c:95  */
void dWorldDestroy(dWorldID);                                  /* src/main.c:266 */
void dWorldSetGravity(dWorldID, dReal x, dReal y, dReal z);    /* src/main.c:96  */
void dWorldGetGravity(dWorldID, dVector3 gravity);
void dWorldSetERP(dWorldID, dReal erp);
dReal dWorldGetERP(dWorldID);
void dWorldSetCFM(dWorldID, dReal cfm);
dReal dWorldGetCFM(dWorldID);
void dWorldSetQuickStepNumIterations(dWorldID, int num);
int dWorldGetQuickStepNumIterations(dWorldID);
void dWorldSetQuickStepW(dWorldID, dReal over_relaxation);
dReal dWorldGetQuickStepW(dWorldID);
void dWorldSetContactMaxCorrectingVel(dWorldID, dReal vel);
dReal dWorldGetContactMaxCorrectingVel(dWorldID);
void dWorldSetContactSurfaceLayer(dWorldID, dReal depth);
dReal dWorldGetContactSurfaceLayer(dWorldID);
/* Both steppers run the same graph-coloured SOR/PGS solver (src/main.c:213 calls dWorldStep;
 * north_star names dWorldQuickStep).  dWorldStep can be told to iterate towards libode's exact LCP answer:
 * dWorldSetStepSolverB200 in ode_b200.h.  Return 1 on success, 0 on failure. */
int dWorldStep(dWorldID, dReal stepsize);                      /* src/main.c:213 */
int dWorldQuickStep(dWorldID, dReal stepsize);

/* ------------------------------------------------------------------ bodies */
dBodyID dBodyCreate(dWorldID);                                 /* src/main.c:703 */
void dBodyDestroy(dBodyID);                                    /* src/main.c:261 */
void dBodySetPosition(dBodyID, dReal x, dReal y, dReal z);     /* src/main.c:708 */
void dBodySetRotation(dBodyID, const dMatrix3 R);              /* src/main.c:709 */
void dBodySetQuaternion(dBodyID, const dQuaternion q);
void dBodySetLinearVel(dBodyID, dReal x, dReal y, dReal z);
void dBodySetAngularVel(dBodyID, dReal x, dReal y, dReal z);
const dReal *dBodyGetPosition(dBodyID);                        /* src/main.c:229 */
const dReal *dBodyGetRotation(dBodyID);                        /* src/main.c:230 */
const dReal *dBodyGetQuaternion(dBodyID);
const dReal *dBodyGetLinearVel(dBodyID);
const dReal *dBodyGetAngularVel(dBodyID);
void dBodySetMass(dBodyID, const dMass *mass);
void dBodyGetMass(dBodyID, dMass *mass);
void dBodySetKinematic(dBodyID);                               /* src/main.c:712 */
void dBodySetDynamic(dBodyID);
int dBodyIsKinematic(dBodyID);
void dBodySetGravityMode(dBodyID, int mode);
int dBodyGetGravityMode(dBodyID);
void dBodySetGyroscopicMode(dBodyID, int enabled);
int dBodyGetGyroscopicMode(dBodyID);
void dBodyAddForce(dBodyID, dReal fx, dReal fy, dReal fz);     /* src/main.c:532 (comment) */
void dBodyAddTorque(dBodyID, dReal fx, dReal fy, dReal fz);
const dReal *dBodyGetForce(dBodyID);
const dReal *dBodyGetTorque(dBodyID);
void dBodySetData(dBodyID, void *data);
void *dBodyGetData(dBodyID);
dWorldID dBodyGetWorld(dBodyID);

/* ------------------------------------------------------------------ mass helpers */
void dMassSetZero(dMass *);
void dMassSetParameters(dMass *, dReal themass, dReal cgx, dReal cgy, dReal cgz, dReal I11,
                        dReal I22, dReal I33, dReal I12, dReal I13, dReal I23);
void dMassSetSphere(dMass *, dReal density, dReal radius);
void dMassSetSphereTotal(dMass *, dReal total_mass, dReal radius);
void dMassSetBox(dMass *, dReal density, dReal lx, dReal ly, dReal lz);
void dMassSetBoxTotal(dMass *, dReal total_mass, dReal lx, dReal ly, dReal lz);
void dMassAdjust(dMass *, dReal newmass);

/* ------------------------------------------------------------------ spaces / geoms */
dSpaceID dHashSpaceCreate(dSpaceID parent);                    /* src/main.c:97  */
dSpaceID dSimpleSpaceCreate(dSpaceID parent);
void dSpaceDestroy(dSpaceID);
int dSpaceGetNumGeoms(dSpaceID);
dGeomID dSpaceGetGeom(dSpaceID, int i);

typedef void dNearCallback(void *data, dGeomID o1, dGeomID o2);
void dSpaceCollide(dSpaceID, void *data, dNearCallback *callback); /* src/main.c:212 */
/* low 16 bits of flags = max contacts; skip = byte stride between dContactGeom outputs */
int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom *contact, int skip); /* :678 */

dGeomID dCreateSphere(dSpaceID, dReal radius);                 /* src/main.c:717 */
dGeomID dCreateBox(dSpaceID, dReal lx, dReal ly, dReal lz);    /* src/main.c:720, :743 */
dGeomID dCreatePlane(dSpaceID, dReal a, dReal b, dReal c, dReal d);
void dGeomDestroy(dGeomID);                                    /* src/main.c:263 */
void dGeomSetBody(dGeomID, dBodyID);                           /* src/main.c:726 */
dBodyID dGeomGetBody(dGeomID);                                 /* src/main.c:691 */
void dGeomSetPosition(dGeomID, dReal x, dReal y, dReal z);     /* src/main.c:748 */
void dGeomSetRotation(dGeomID, const dMatrix3 R);              /* src/main.c:749 */
void dGeomSetQuaternion(dGeomID, const dQuaternion q);
const dReal *dGeomGetPosition(dGeomID);                        /* src/main.c:232 */
const dReal *dGeomGetRotation(dGeomID);                        /* src/main.c:233 */
void dGeomGetQuaternion(dGeomID, dQuaternion result);
void dGeomGetAABB(dGeomID, dReal aabb[6]);
int dGeomGetClass(dGeomID);
void dGeomSetCategoryBits(dGeomID, unsigned long bits);        /* src/main.c:724, :751, :752 */
void dGeomSetCollideBits(dGeomID, unsigned long bits);         /* src/main.c:725 */
unsigned long dGeomGetCategoryBits(dGeomID);
unsigned long dGeomGetCollideBits(dGeomID);
void dGeomSetData(dGeomID, void *data);
void *dGeomGetData(dGeomID);
dSpaceID dGeomGetSpace(dGeomID);
dReal dGeomSphereGetRadius(dGeomID);
void dGeomSphereSetRadius(dGeomID, dReal radius);
void dGeomBoxGetLengths(dGeomID, dVector3 result);
void dGeomBoxSetLengths(dGeomID, dReal lx, dReal ly, dReal lz);
void dGeomPlaneGetParams(dGeomID, dVector4 result);
void dGeomPlaneSetParams(dGeomID, dReal a, dReal b, dReal c, dReal d);

/* trimesh (north_star: sphere-vs-trimesh, teapot.obj) */
typedef int dTriCallback(dGeomID TriMesh, dGeomID RefObject, int TriangleIndex);
typedef void dTriArrayCallback(dGeomID TriMesh, dGeomID RefObject, const int *TriIndices,
                               int TriCount);
typedef int dTriRayCallback(dGeomID TriMesh, dGeomID Ray, int TriangleIndex, dReal u, dReal v);
dTriMeshDataID dGeomTriMeshDataCreate(void);
void dGeomTriMeshDataDestroy(dTriMeshDataID);
/* Vertices: float[3] at VertexStride bytes; Indices: dTriIndex[3] at TriStride bytes;
 * IndexCount = 3 * number of triangles.  Data is copied. */
void dGeomTriMeshDataBuildSingle(dTriMeshDataID, const void *Vertices, int VertexStride,
                                 int VertexCount, const void *Indices, int IndexCount,
                                 int TriStride);
dGeomID dCreateTriMesh(dSpaceID, dTriMeshDataID, dTriCallback *, dTriArrayCallback *,
                       dTriRayCallback *);

/* ------------------------------------------------------------------ contact joints */
dJointGroupID dJointGroupCreate(int max_size);                 /* src/main.c:98  */
void dJointGroupEmpty(dJointGroupID);                          /* src/main.c:214 */
void dJointGroupDestroy(dJointGroupID);                        /* src/main.c:265 */
dJointID dJointCreateContact(dWorldID, dJointGroupID, const dContact *); /* src/main.c:690 */
void dJointAttach(dJointID, dBodyID body1, dBodyID body2);     /* src/main.c:691 */
dBodyID dJointGetBody(dJointID, int index);

/* ------------------------------------------------------------------ rotation helpers */
void dRSetIdentity(dMatrix3 R);
void dRFromAxisAndAngle(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal angle);
void dRFromEulerAngles(dMatrix3 R, dReal phi, dReal theta, dReal psi);
void dQSetIdentity(dQuaternion q);
void dQFromAxisAndAngle(dQuaternion q, dReal ax, dReal ay, dReal az, dReal angle);
void dQtoR(const dQuaternion q, dMatrix3 R);
void dRtoQ(const dMatrix3 R, dQuaternion q);
void dPlaneSpace(const dVector3 n, dVector3 p, dVector3 q);

#ifdef __cplusplus
}
#endif

#endif
