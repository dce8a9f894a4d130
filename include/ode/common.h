/* ode/common.h -- libode_b200's replacement for the header the reference includes at
 * inc/body.h:4 (`#include "ode/common.h"`).  Only the types the reference's host code
 * touches are declared: dReal, the fixed-size vector/matrix typedefs and the opaque IDs
 * (reference uses dBodyID / dGeomID in inc/body.h:20-24).
 *
 * dReal is float: the engine is fp32 end to end ("dReal=float").
 */
#ifndef ODE_B200_COMMON_H
#define ODE_B200_COMMON_H

#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef dSINGLE
#define dSINGLE 1
#endif

typedef float dReal;

#ifdef INFINITY
#define dInfinity ((dReal)INFINITY)
#else
#define dInfinity ((dReal)(1.0 / 0.0))
#endif

#define REAL(x) (x##f)

/* ODE pads 3-vectors to 4 and 3x3 matrices to 3 rows of 4 (R[i*4+j]). */
typedef dReal dVector3[4];
typedef dReal dVector4[4];
typedef dReal dMatrix3[4 * 3];
typedef dReal dMatrix4[4 * 4];
typedef dReal dQuaternion[4]; /* (w, x, y, z) */

struct dxWorld;
struct dxSpace;
struct dxBody;
struct dxGeom;
struct dxJoint;
struct dxJointGroup;
struct dxTriMeshData;

typedef struct dxWorld *dWorldID;
typedef struct dxSpace *dSpaceID;
typedef struct dxBody *dBodyID;
typedef struct dxGeom *dGeomID;
typedef struct dxJoint *dJointID;
typedef struct dxJointGroup *dJointGroupID;
typedef struct dxTriMeshData *dTriMeshDataID;

typedef unsigned int dTriIndex;

#ifdef __cplusplus
}
#endif

#endif
