// Entry points of include/m3d.h that are not implemented yet in this build.
// They fail loudly with M3D_ERR_UNSUPPORTED; nothing here computes on the CPU.
#include "api_common.h"

using namespace m3d;

#define M3D_TODO(name) return fail(M3D_ERR_UNSUPPORTED, name ": not implemented in this build")

extern "C" {

#ifndef M3D_HAVE_SCENE
int32_t m3d_scene_builder_create(m3d_ctx *, m3d_scene_builder **) { M3D_TODO("m3d_scene_builder_create"); }
void m3d_scene_builder_destroy(m3d_scene_builder *) {}
int32_t m3d_scene_add_material(m3d_scene_builder *, const m3d_material_desc *, int32_t *) { M3D_TODO("m3d_scene_add_material"); }
int32_t m3d_scene_add_mesh(m3d_scene_builder *, const float *, int64_t, const float *, int32_t, uint32_t, const m3d_transform *, int32_t *) { M3D_TODO("m3d_scene_add_mesh"); }
int32_t m3d_scene_add_sphere(m3d_scene_builder *, const double *, double, int32_t, uint32_t, const m3d_transform *, int32_t *) { M3D_TODO("m3d_scene_add_sphere"); }
int32_t m3d_scene_add_rect(m3d_scene_builder *, const double *, const double *, int32_t, uint32_t, const m3d_transform *, int32_t *) { M3D_TODO("m3d_scene_add_rect"); }
int32_t m3d_scene_add_cylinder(m3d_scene_builder *, const double *, const double *, double, int32_t, uint32_t, const m3d_transform *, int32_t *) { M3D_TODO("m3d_scene_add_cylinder"); }
int32_t m3d_scene_build(m3d_scene_builder *, uint32_t, m3d_scene **) { M3D_TODO("m3d_scene_build"); }
void m3d_scene_destroy(m3d_scene *) {}
int32_t m3d_scene_bounds(const m3d_scene *, double *, double *) { M3D_TODO("m3d_scene_bounds"); }
int32_t m3d_scene_cast(m3d_scene *, const float *, const float *, int64_t, float *, int32_t *, int32_t *, float *, uint32_t, m3d_stats *) { M3D_TODO("m3d_scene_cast"); }
#endif

#ifndef M3D_HAVE_RAYCAST
int32_t m3d_render_raycast(m3d_scene *, const m3d_camera *, const m3d_point_light *, int32_t, int32_t, int32_t, const m3d_partition *, float *, m3d_stats *) { M3D_TODO("m3d_render_raycast"); }
int32_t m3d_render_raycast_device(m3d_scene *, const m3d_camera *, const m3d_point_light *, int32_t, int32_t, int32_t, const m3d_partition *, void *, void *, m3d_stats *) { M3D_TODO("m3d_render_raycast_device"); }
int32_t m3d_finalize_image_device(m3d_ctx *, const void *, int64_t, double, void *, void *, void *) { M3D_TODO("m3d_finalize_image_device"); }
#endif

#ifndef M3D_HAVE_PATH
int32_t m3d_render_path(m3d_scene *, const m3d_camera *, const m3d_point_light *, int32_t, const m3d_path_params *, int32_t, int32_t, const m3d_partition *, int32_t, float *, float *, m3d_stats *) { M3D_TODO("m3d_render_path"); }
int32_t m3d_render_path_device(m3d_scene *, const m3d_camera *, const m3d_point_light *, int32_t, const m3d_path_params *, int32_t, int32_t, const m3d_partition *, int32_t, void *, void *, void *, m3d_stats *) { M3D_TODO("m3d_render_path_device"); }
#endif

#ifndef M3D_HAVE_BIDIR
int32_t m3d_render_bidir(m3d_scene *, const m3d_camera *, const m3d_area_light *, int32_t,
                         const m3d_bidir_params *, int32_t, int32_t, const m3d_partition *, int32_t,
                         float *, float *, m3d_stats *) {
  M3D_TODO("m3d_render_bidir");
}
int32_t m3d_render_bidir_device(m3d_scene *, const m3d_camera *, const m3d_area_light *, int32_t,
                                const m3d_bidir_params *, int32_t, int32_t, const m3d_partition *,
                                int32_t, void *, void *, void *, m3d_stats *) {
  M3D_TODO("m3d_render_bidir_device");
}
#endif

}  // extern "C"
