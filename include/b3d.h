/*
 * libb3d — C ABI of the B200-native hot path of NikolaZubic/2dimageto3dmodel.
 *
 * The reference has no FFI of its own: its boundary is the Python call surface
 * (SURVEY.md §8b).  Every entry point below is what a reference-side binding for
 * that call would wrap; the reference function it replaces is cited as
 * /root/reference/code/<file>:<lines>.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions (all entry points)
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory,
 *     contiguous, fp32 unless stated, base pointers 16-byte aligned;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never
 *     allocates or frees persistent device memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no
 *     internal synchronisation;
 *   - return 0 on success, a negative B3D_E* code otherwise; b3d_last_error()
 *     returns a thread-local message.  Nothing throws or exits across the ABI;
 *   - stateless and re-entrant.
 */
#ifndef B3D_H_
#define B3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B3D_API __attribute__((visibility("default")))
#else
#define B3D_API
#endif

#define B3D_OK 0
#define B3D_EINVAL (-1)  /* bad argument (null pointer, negative size, unsupported shape) */
#define B3D_EALIGN (-2)  /* pointer not 16-byte aligned */
#define B3D_ECUDA (-3)   /* CUDA runtime error (message in b3d_last_error) */

/* semantics switch of the point-cloud path, SURVEY.md App. A */
#define B3D_MODE_REFERENCE 0 /* "R": the reference as written (quirks D3,D4,D5,D10 kept) */
#define B3D_MODE_PAPER 1     /* "P": paper-intended */

B3D_API const char* b3d_last_error(void);
B3D_API int b3d_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
B3D_API uint64_t b3d_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Point-cloud "effective loss" path
 *   EffectiveLossFunction.forward            utils/effective_loss_function.py:58-81
 * ------------------------------------------------------------------------------------------ */

/* Rotate by the normalised quaternion, perspective-project, map to grid coordinates.
 *   PointsQuaternionsRotator.rotate_points   quaternions/points_quaternions.py:41-81
 *   QuaternionOperations.quaternion_multiplication / _conjugate   quaternions/operations.py:68-97,120-136
 *   CameraUtilities.transformation_3d_coord_to_camera_coord       camera/coordinate_system_transformation.py:20-39
 *   TrilinearInterpolation.get_grid / get_point_cloud_object_borders  utils/trilinear_interpolation.py:17-35
 * points [B,N,3] (columns z,y,x), quat [B,4] (w,x,y,z); fov / cam_dist are the reference's
 * field_of_view = 1.875 and camera_view_distance = 2.0 (effective_loss_function.py:69-70).
 * pg     [B,N,4]  out: grid coords (gz,gy,gx) = (V-1)(c+0.5) and flag 1.0/0.0 = in bounds
 * coords [B,N,3]  out, nullable: camera coords c
 * base   [B,N,3]  out, nullable, int32: floor(g) — the reference's index buffer for corner (0,0,0)
 *                 (trilinear_interpolation.py:47-52); bit-exact target
 * inb    [B,N]    out, nullable, uint8: in-bounds mask (bit-exact target)                      */
B3D_API int b3d_pc_project(const float* points, const float* quat, int B, int N, int V, float fov,
                   float cam_dist, float* pg, float* coords, int32_t* base, uint8_t* inb,
                   void* stream);

/* Splat + z-blur + scale/clamp + ray termination + silhouette, one fused kernel (mode R) or
 * splat / separable-blur / ray-march kernels over a materialised grid (mode P).
 *   TrilinearInterpolation.trilinear_interpolation / positions_update  utils/trilinear_interpolation.py:37-74
 *   VoxelsSmooth.smooth                                                utils/smooth_voxels.py:44-84
 *   EffectiveLossFunction.termination_probs + sum + flip               utils/effective_loss_function.py:18-56,79-81
 * pg [B,N,4] from b3d_pc_project; taps [ktaps] the 1-D smoothing kernel (host computes it with the
 * reference's expression, smooth_voxels.py:24-31); scale [B] nullable; sil [B,V,V] out.
 * workspace: b3d_pc_silhouette_workspace_bytes(B,V,mode) bytes (0 for mode R; may be NULL then). */
B3D_API size_t b3d_pc_silhouette_workspace_bytes(int B, int V, int mode);
B3D_API int b3d_pc_silhouette_fwd(const float* pg, const float* taps, int ktaps, const float* scale, int B,
                          int N, int V, int mode, float* sil, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Same, with the taps in HOST memory (saves the device->host read of 21 floats; this is the entry the
 * Python wrapper uses, its taps are computed on the CPU with the reference's torch expression). */
B3D_API int b3d_pc_silhouette_fwd_hosttaps(const float* pg, const float* taps_host, int ktaps,
                                   const float* scale, int B, int N, int V, int mode, float* sil,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above: dsil [B,V,V] -> dpg [B,N,4] (d/d grid coords, .w unused), dscale [B]
 * (nullable iff scale is NULL; zeroed by the call). */
B3D_API int b3d_pc_silhouette_bwd(const float* pg, const float* taps, int ktaps, const float* scale,
                          const float* dsil, int B, int N, int V, int mode, float* dpg,
                          float* dscale, void* workspace, size_t workspace_bytes, void* stream);

B3D_API int b3d_pc_silhouette_bwd_hosttaps(const float* pg, const float* taps_host, int ktaps,
                                   const float* scale, const float* dsil, int B, int N, int V,
                                   int mode, float* dpg, float* dscale, void* workspace,
                                   size_t workspace_bytes, void* stream);

/* Backward of b3d_pc_project: (pg, dpg) [B,N,4] -> dpoints [B,N,3], dquat [B,4] (zeroed by the call). */
B3D_API int b3d_pc_project_bwd(const float* points, const float* quat, const float* pg, const float* dpg,
                       int B, int N, int V, float fov, float cam_dist, float* dpoints,
                       float* dquat, void* stream);

/* Materialised occupancy grid [B,V,V,V] (clamped to [0,1]) — the tensor
 * TrilinearInterpolation.trilinear_interpolation returns (trilinear_interpolation.py:74).
 * Used by mode P and by the parity tests; grid is zeroed by the call. */
B3D_API int b3d_pc_splat_grid(const float* pg, int B, int N, int V, int mode, float* grid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B3D_H_ */
