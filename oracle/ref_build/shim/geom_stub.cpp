/* Stand-ins for the few meep_geom symbols that the stepping core links against
 * (vec.cpp grid_volume::get_cost, structure.cpp choose_chunkdivision).  The real
 * definitions live in the reference's meepgeom.cpp, which needs libctlgeom (absent).
 * With resolution == 0 / split_chunks_evenly == true the reference takes its
 * "split by voxel count" branch, i.e. the cost model is never consulted.
 * TEST/BASELINE INFRASTRUCTURE ONLY. */
#include "meepgeom.hpp"

namespace meep_geom {

double fragment_stats::tol = 0;
int fragment_stats::maxeval = 0;
int fragment_stats::resolution = 0;
meep::ndim fragment_stats::dims = meep::D1;
geometric_object_list fragment_stats::geom = {0, 0};
std::vector<dft_data> fragment_stats::dft_data_list;
std::vector<meep::volume> fragment_stats::pml_1d_vols;
std::vector<meep::volume> fragment_stats::pml_2d_vols;
std::vector<meep::volume> fragment_stats::pml_3d_vols;
std::vector<meep::volume> fragment_stats::absorber_vols;
material_type_list fragment_stats::extra_materials = material_type_list();
bool fragment_stats::split_chunks_evenly = true;
bool fragment_stats::eps_averaging = false;

fragment_stats::fragment_stats(geom_box &bx)
    : num_anisotropic_eps_pixels(0), num_anisotropic_mu_pixels(0), num_nonlinear_pixels(0),
      num_susceptibility_pixels(0), num_nonzero_conductivity_pixels(0), num_1d_pml_pixels(0),
      num_2d_pml_pixels(0), num_3d_pml_pixels(0), num_dft_pixels(0), num_pixels_in_box(0),
      box(bx) {}

void fragment_stats::compute() {}

double fragment_stats::cost() const { return 1.0; }

geom_box gv2box(const meep::volume &v) {
  geom_box b;
  meep::vec lo = v.get_min_corner(), hi = v.get_max_corner();
  b.low.x = b.low.y = b.low.z = b.high.x = b.high.y = b.high.z = 0;
  LOOP_OVER_DIRECTIONS(v.dim, d) {
    double l = lo.in_direction(d), h = hi.in_direction(d);
    switch (d) {
      case meep::X: case meep::R: b.low.x = l; b.high.x = h; break;
      case meep::Y: case meep::P: b.low.y = l; b.high.y = h; break;
      case meep::Z: b.low.z = l; b.high.z = h; break;
      default: break;
    }
  }
  return b;
}

} // namespace meep_geom

namespace meep_geom {
// defined in the reference's material_data.cpp (not built: it needs libctl's vector3 helpers)
material_type_list::material_type_list() : items(NULL), num_items(0) {}
} // namespace meep_geom
