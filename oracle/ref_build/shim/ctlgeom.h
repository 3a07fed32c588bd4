/* Minimal stand-in for libctl's <ctlgeom.h> (libctl/libctlgeom is a third-party
 * dependency that is not installed in this image).  Only the plain-data types that
 * the reference's meepgeom.hpp / material_data.hpp mention by value are declared,
 * so that vec.cpp and structure.cpp (which include meepgeom.hpp) compile.  No
 * geometry functionality exists in this build.  TEST/BASELINE INFRASTRUCTURE ONLY. */
#ifndef MEEP_B200_SHIM_CTLGEOM_H
#define MEEP_B200_SHIM_CTLGEOM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef double number;
typedef int integer;
typedef short boolean;
typedef struct { number x, y, z; } vector3;
typedef struct { number re, im; } cnumber;
typedef struct { cnumber x, y, z; } cvector3;
typedef struct { vector3 c0, c1, c2; } matrix3x3;
typedef struct { vector3 low, high; } geom_box;
typedef struct geom_box_tree_struct *geom_box_tree;
typedef struct geometric_object_struct {
  void *material;
  vector3 center;
  int which_subclass;
  void *subclass_data;
} geometric_object;
typedef struct {
  int num_items;
  geometric_object *items;
} geometric_object_list;
#ifdef __cplusplus
}
#endif
#endif
