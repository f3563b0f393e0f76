// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for UFC 2.0.5 <ufc.h>.
//
// The reference's FFC-generated element kernels (comri/*/hpc-fenics-cpp/ufc/*.cpp,
// `*_integral_*::tabulate_tensor`) only touch `ufc::cell::coordinates`.  This shim
// declares just enough of the ufc namespace for those function bodies to compile
// when oracle/build_ref.py lifts them, at build time, out of /root/reference into a
// scratch translation unit.  Written from the UFC 2.0.5 interface description, not
// copied from it.
#ifndef BTFEM_ORACLE_UFC_SHIM_H
#define BTFEM_ORACLE_UFC_SHIM_H

namespace ufc {

enum shape { interval, triangle, quadrilateral, tetrahedron, hexahedron };

struct cell {
  shape cell_shape;
  unsigned int topological_dimension;
  unsigned int geometric_dimension;
  unsigned int** entity_indices;
  double** coordinates;   // coordinates[vertex][xyz]
  int index;
  int local_facet;
  int mesh_identifier;
  cell()
      : cell_shape(tetrahedron), topological_dimension(3), geometric_dimension(3),
        entity_indices(0), coordinates(0), index(0), local_facet(-1), mesh_identifier(-1) {}
};

struct cell_integral { virtual ~cell_integral() {} };
struct exterior_facet_integral { virtual ~exterior_facet_integral() {} };
struct interior_facet_integral { virtual ~interior_facet_integral() {} };

}  // namespace ufc

#endif
