/* svslam.h — dependency-free C ABI of the B200-native StereoVision-SLAM hot path.
 *
 * This is the drop-in boundary: each entry point replaces what
 * one third-party call site inside the reference's private Frontend / Backend / DenseReconstruction
 * methods computes.  Plain pointers and sizes only; every function returns 0 (SVS_OK) or a
 * negative svs_status and never throws; svs_last_error() returns the message of the last failure
 * on that context.  All pointers are caller-owned HOST memory unless the parameter name ends in
 * `_dev`.  A context is single-threaded; distinct contexts are independent (the reference's
 * frontend thread and backend thread each own one) and each owns one CUDA stream.
 *
 * There is NO CPU fallback: without a CUDA device svs_create() fails.
 *
 * Conventions
 *   pose / extrinsic  double[7] = qx qy qz qw tx ty tz   (Sophus::SE3d: unit quaternion + translation, T_cw)
 *   intrinsics        double[4] = fx fy cx cy            (Camera::K(), src/camera.cpp:14-21)
 *   pixel coordinates float (x, y) interleaved
 *   "ragged" batches  CSR offsets: item i of problem b lives at [off[b], off[b+1])
 */
#ifndef SVSLAM_H
#define SVSLAM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVS_API __attribute__((visibility("default")))

typedef enum {
    SVS_OK = 0,
    SVS_ERR_CUDA = -1,       /* CUDA runtime error (message has the cudaError string) */
    SVS_ERR_ARG = -2,        /* invalid argument */
    SVS_ERR_CAPACITY = -3,   /* an internal capacity was exceeded (message says which) */
    SVS_ERR_NODEVICE = -4    /* no CUDA device / wrong architecture */
} svs_status;

typedef struct svs_ctx svs_ctx;
typedef struct svs_frameset svs_frameset;

/* ---------------------------------------------------------------- context */
SVS_API svs_ctx *svs_create(int device);                 /* NULL on failure (see svs_create_error) */
SVS_API const char *svs_create_error(void);
SVS_API void svs_destroy(svs_ctx *ctx);
SVS_API const char *svs_last_error(svs_ctx *ctx);
SVS_API int svs_version(void);
SVS_API int svs_sync(svs_ctx *ctx);                       /* wait for the context stream (by the wait mode below) */
/* Host wait policy of every blocking call on this context: 0 (default) = cudaStreamSynchronize, the driver spins on the
 * stream (lowest wake-up latency, one busy core per waiting host thread); 1 = the thread sleeps on a blocking-sync event
 * (for hosts with fewer cores than waiting threads, e.g. 8 ranks x 2 contexts on a 32-core box).  Results are identical. */
SVS_API int svs_set_wait_mode(svs_ctx *ctx, int mode);
/* Scheduling of the window solver behind svs_ba_optimize (no effect on results).  high_priority 1: the solver runs on a
 * high-priority stream of its own, so that its few long CTAs are placed ahead of the queued CTAs of machine-filling kernels
 * launched by other contexts; threads_per_window 0 = automatic, 256 or 512 = force that CTA size (256 leaves half of an
 * SM's registers to kernels of other contexts). */
SVS_API int svs_set_ba_schedule(svs_ctx *ctx, int high_priority, int threads_per_window);
/* Diagnostic: how many times an internal grow-only device / pinned buffer was (re)allocated in this process so far.  A
 * regrowth synchronises the device; a primed steady-state step does none. */
SVS_API long long svs_buffer_regrowths(void);
/* Grow every variable-size scratch buffer of the context (and of the frame sets / pipelines created on it) to
 * factor x the largest size requested so far; call once after warm-up, when no call is in flight on the context.  Buffer
 * contents are kept.  factor in [1, 64]. */
SVS_API int svs_reserve_headroom(svs_ctx *ctx, double factor);
SVS_API void *svs_stream(svs_ctx *ctx);                   /* the cudaStream_t, for event timing */
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
SVS_API long long svs_launch_count(svs_ctx *ctx);
/* Optional per-kernel device timing: when enabled every kernel launch of this context is bracketed by CUDA events
 * on the context stream; svs_kernel_timing_get returns accumulated milliseconds and launch counts per kernel class
 * (names via svs_kernel_name) and the number of classes. */
SVS_API int svs_kernel_timing_enable(svs_ctx *ctx, int on);
SVS_API int svs_kernel_timing_reset(svs_ctx *ctx);
SVS_API int svs_kernel_timing_get(svs_ctx *ctx, double *ms, long long *count, int n);
SVS_API const char *svs_kernel_name(int kid);
/* pinned host memory helpers (for callers without their own pinned allocator) */
SVS_API void *svs_host_alloc(size_t bytes);
SVS_API void svs_host_free(void *p);

/* ---------------------------------------------------------------- frame sets
 * A frame set holds, on the device, for each of n_streams independent stereo streams:
 * the current and the previous LEFT image pyramid and the current RIGHT image pyramid
 * (level 0 = the processed image).  Replaces Dataset::NextFrame's two cv::resize calls
 * (src/dataset.cpp:126-134) and the pyramids cv::calcOpticalFlowPyrLK builds internally
 * (src/frontend.cpp:105, :353).
 *   half != 0 : processed image = 0.5x nearest of the input, size (rint(w/2), rint(h/2))
 *   half == 0 : processed image = input
 */
SVS_API svs_frameset *svs_frameset_create(svs_ctx *ctx, int n_streams, int in_w, int in_h, int half,
                                          int lk_win, int lk_max_level);
SVS_API void svs_frameset_destroy(svs_ctx *ctx, svs_frameset *fs);
SVS_API int svs_frameset_size(const svs_frameset *fs, int *w, int *h, int *n_levels);
/* Push one new stereo pair for every stream: previous <- current, then resize + pyramids.
 * left/right: n_streams images, image b at ptr + b*img_stride_bytes, rows row_stride bytes apart.
 * on_device != 0: the pointers are device pointers (inputs already resident in HBM). */
SVS_API int svs_frameset_push(svs_ctx *ctx, svs_frameset *fs, const uint8_t *left, const uint8_t *right,
                              size_t row_stride, size_t img_stride_bytes, int on_device);
/* Same, with one pointer per stream (left[b], right[b]): dense rows `row_stride` apart.
 *   on_device = 0  host memory (pinned for full PCIe bandwidth): staged with strided DMA copies of the rows the resize reads
 *   on_device = 1  device memory
 *   on_device = 2  pinned, device-addressable host memory (cudaHostAlloc / svs_host_alloc): zero-copy — the resize kernel
 *                  reads the frames directly over PCIe, each needed row exactly once */
SVS_API int svs_frameset_push_ptrs(svs_ctx *ctx, svs_frameset *fs, const uint8_t *const *left, const uint8_t *const *right,
                                   size_t row_stride, int on_device);
/* Lazy right-eye ingest.  The frontend reads the right image only when a stream inserts a keyframe
 * (Frontend::FindFeaturesInRight, src/frontend.cpp:72-141), i.e. for a few percent of the frames: svs_frameset_push_ptrs and
 * svs_frameset_prefetch_ptrs accept right == NULL (left eye only), and this call ingests the CURRENT right image
 * (resize + pyramids) of the selected streams only: right[k] belongs to stream stream_ids[k]. */
SVS_API int svs_frameset_fetch_right_ptrs(svs_ctx *ctx, svs_frameset *fs, const int32_t *stream_ids, int n_sel,
                                          const uint8_t *const *right, size_t row_stride, int on_device);
/* Image bytes this frame set has read from HOST memory so far (bench.py's h2d_bytes_per_step). */
SVS_API long long svs_frameset_h2d_bytes(const svs_frameset *fs);
/* Asynchronous ingest of the NEXT stereo pair (double buffering): starts the resize + pyramids of the pair the caller will
 * push next into spare buffers on the context's second (ingest) stream and returns immediately, so that the PCIe
 * transfer of frame t+1 overlaps the tracking / optimisation kernels of frame t.  The following
 * svs_frameset_push_ptrs with the SAME pointers, row stride and mode only rotates buffers; a push with anything else
 * discards the prefetch and ingests normally.  The frames must stay valid and unchanged until that push. */
SVS_API int svs_frameset_prefetch_ptrs(svs_ctx *ctx, svs_frameset *fs, const uint8_t *const *left, const uint8_t *const *right,
                                       size_t row_stride, int on_device);
/* Copy a pyramid level back (tests).  which: 0 = current left, 1 = previous left, 2 = current right */
SVS_API int svs_frameset_download(svs_ctx *ctx, svs_frameset *fs, int stream, int which, int level,
                                  uint8_t *out, int out_stride);

/* ---------------------------------------------------------------- a0 : half-resolution resize
 * Replaces cv::resize(src, dst, Size(), 0.5, 0.5, INTER_NEAREST) at src/dataset.cpp:128-129. */
SVS_API int svs_half_nearest(svs_ctx *ctx, const uint8_t *src, int w, int h, int stride, int n_images,
                             size_t img_stride_bytes, uint8_t *dst /* n * rint(h/2) * rint(w/2) */);

/* ---------------------------------------------------------------- a1 : GFTT
 * Replaces kp_detector_->detect(img, keypoints, mask) at src/frontend.cpp:51 (detector created at :24 as
 * cv::GFTTDetector::create(num_features, 0.01, 20)) together with the mask Frontend::DetectFeatures
 * draws at :42-47.  The mask is given either as an image (mask != NULL, 0 = excluded) or as the list of
 * box centres (occupied_xy: the positions of the existing left features; the excluded box is
 * [rint(x-10), rint(x+10)] x [rint(y-10), rint(y+10)], f32 subtraction, round-half-even).
 * oracle_simd_granule: SIMD granule of the OpenCV build being matched (its Sobel-dy row filter uses a
 * fused multiply-add for columns < g*floor(w/g) and separate mul+add for the tail; 32 = AVX-512 build,
 * 0 = unfused everywhere).  Output: strongest-first (x, y) with integer values and the corner response. */
SVS_API int svs_gftt_detect(svs_ctx *ctx, const uint8_t *img, int w, int h, int stride,
                            const uint8_t *mask, int mask_stride,
                            const float *occupied_xy, int n_occupied,
                            int max_corners, double quality, double min_distance, int oracle_simd_granule,
                            float *out_xy /* 2*max_corners */, float *out_response /* max_corners */, int *out_n);
/* The corner-response map alone (cv::cornerMinEigenVal(img, 3, 3)); tests and profiling. */
SVS_API int svs_corner_min_eig(svs_ctx *ctx, const uint8_t *img, int w, int h, int stride,
                               int oracle_simd_granule, float *out /* h*w */);
/* Batched over the CURRENT LEFT images of selected streams of a frame set. */
SVS_API int svs_gftt_detect_batch(svs_ctx *ctx, svs_frameset *fs, const int32_t *stream_ids, int n_sel,
                                  const int32_t *occ_off /* n_sel+1 */, const float *occupied_xy,
                                  int max_corners, double quality, double min_distance, int oracle_simd_granule,
                                  float *out_xy /* n_sel*2*max */, float *out_response /* n_sel*max */,
                                  int32_t *out_n /* n_sel */);

/* ---------------------------------------------------------------- a2 / a3 : pyramidal LK
 * Replaces cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win,win), max_level,
 * TermCriteria(COUNT+EPS, max_iter, eps), OPTFLOW_USE_INITIAL_FLOW) at src/frontend.cpp:105-109 (left ->
 * right) and :353-357 (last -> current).  next_xy is in/out (initial flow in, result out); err is not
 * produced (the reference never reads it) but its side effect on status is. */
SVS_API int svs_lk_track(svs_ctx *ctx, const uint8_t *prev, const uint8_t *next, int w, int h, int stride,
                         const float *prev_xy, float *next_xy, int n, int win, int max_level, int max_iter,
                         double eps, uint8_t *status);
/* Batched over the streams of a frame set.  pair: 0 = previous left -> current left (TrackLastFrame),
 * 1 = current left -> current right (FindFeaturesInRight).  Points of stream b: [off[b], off[b+1]). */
SVS_API int svs_lk_track_batch(svs_ctx *ctx, svs_frameset *fs, int pair, const int32_t *off /* n_streams+1 */,
                               const float *prev_xy, float *next_xy, int max_iter, double eps, uint8_t *status);

/* ---------------------------------------------------------------- a4 : triangulation
 * Replaces slam::triangulation(poses, points, pworld) (include/StereoVisionSLAM/algorithm.h:10-87) as called
 * at src/frontend.cpp:174 and :286 with Camera::pixel2camera (src/camera.cpp:58-72) for a rectified pair:
 * left extrinsic identity, right extrinsic [I | (-baseline, 0, 0)].  out_ok = sigma4/sigma3 < 1e-2. */
SVS_API int svs_triangulate(svs_ctx *ctx, const float *left_xy, const float *right_xy, int n,
                            const double K_left[4], const double K_right[4], double baseline,
                            double *out_xyz /* 3n */, uint8_t *out_ok /* n */);

/* ---------------------------------------------------------------- a5 : pose-only LM
 * Replaces the g2o block of Frontend::EstimateCurrentPose (src/frontend.cpp:408-527): one VertexPose,
 * m EdgeProjectionPoseOnly (g2o_types.h:94-174), Huber(1.0), `rounds` x optimize(`iters`) with the
 * chi2 > chi2_th outlier re-classification between rounds and the robust kernel dropped after round
 * index 2.  Batched: problem b owns edges [off[b], off[b+1]). */
typedef struct {
    int32_t iterations, trials, linearizations, solves;
    double lambda, chi2;
} svs_lm_stats;
SVS_API int svs_pose_only_lm(svs_ctx *ctx, int n_prob, const int32_t *off, const double *pts_w /* 3*M */,
                             const double *uv /* 2*M */, const double *K /* 4*n_prob */,
                             const double *T0 /* 7*n_prob */, double chi2_th, int rounds, int iters,
                             double *T_out /* 7*n_prob */, uint8_t *outlier_out /* M */,
                             int32_t *n_inlier /* n_prob */, svs_lm_stats *stats /* n_prob or NULL */);

/* ---------------------------------------------------------------- a7 : bundle adjustment
 * Replaces the g2o block of Backend::Optimize (src/backend.cpp:22-164): VertexPose per keyframe,
 * marginalised VertexXYZ per landmark, EdgeProjection (g2o_types.h:176-229) per observation with
 * Huber(huber_delta), Levenberg-Marquardt with Schur complement and a dense pivoted LDLT of the
 * reduced camera system, max_iter iterations, no vertex fixed.  The chi2 post-pass
 * (src/backend.cpp:167-213) stays with the caller and uses edge_chi2_out.
 * jacobian_mode: 0 analytic, 1 central differences with delta = 1e-9 (g2o's default for this edge).
 * Batched: problem b owns keyframes [kf_off[b],kf_off[b+1]), landmarks [lm_off..), edges [e_off..);
 * edge_kf / edge_lm are indices LOCAL to the problem.  Keyframes/landmarks in ascending id order. */
typedef struct {
    int32_t iterations, trials, linearizations, solves;
    double lambda, chi2, chi2_init;
} svs_ba_stats;
SVS_API int svs_ba_optimize(svs_ctx *ctx, int n_prob, const int32_t *kf_off, double *poses /* 7*sumN in/out */,
                            const int32_t *lm_off, double *lms /* 3*sumL in/out */,
                            const int32_t *e_off, const int32_t *edge_kf, const int32_t *edge_lm,
                            const uint8_t *edge_cam /* 0 left, 1 right */, const double *edge_uv,
                            const double K_left[4], const double K_right[4],
                            const double ext_left[7], const double ext_right[7],
                            double huber_delta, int max_iter, int jacobian_mode,
                            double *edge_chi2_out /* sumE */, svs_ba_stats *stats /* n_prob or NULL */);

/* Wall time svs_ba_optimize has spent on this context in: building the problem structure on the host | packing + enqueueing |
 * waiting for the device and unpacking (diagnostics for bench.py). */
SVS_API int svs_ba_host_seconds(svs_ctx *ctx, double out[3]);

/* ---------------------------------------------------------------- a7, sharded : large / multi-GPU bundle adjustment
 * Same problem and solver as svs_ba_optimize — the g2o block of Backend::Optimize, src/backend.cpp:22-164, ending in
 * optimizer.optimize(max_iter) at :163-164 — for windows too large for one CTA (config 4: N = 50, L = 1e5) and for landmark
 * sharding across GPUs: every shard holds ALL n_kf poses and a disjoint subset of the landmarks with their
 * edges (edge_lm indexes the shard's own landmarks).
 *
 * svs_ba_shard_optimize runs the WHOLE Levenberg-Marquardt loop in one persistent cooperative kernel: accept / reject,
 * lambda schedule and stopping tests are on the device, no host round trip.  With several shards the kernels of all ranks
 * exchange their partial reduced camera systems THEMSELVES through peer memory (no NCCL): every shard owns an exchange
 * window in device memory (svs_ba_shard_window); the caller maps every peer's window into this process (svs_ipc_export /
 * svs_ipc_import over any bootstrap: MPI, sockets, torch.distributed; plain pointers inside one process) and hands the
 * list to svs_ba_shard_set_peers.  Then ALL ranks call svs_ba_shard_optimize (or _launch + _finish) concurrently; each
 * returns the same statistics, poses (replicated) and its own landmarks.  A rank whose peer never arrives fails with
 * SVS_ERR_CUDA after a timeout instead of hanging. */
typedef struct svs_ba_shard svs_ba_shard;
SVS_API svs_ba_shard *svs_ba_shard_create(svs_ctx *ctx, int n_kf, const double *poses, int n_lm, const double *lms, int n_edge,
                                          const int32_t *edge_kf, const int32_t *edge_lm, const uint8_t *edge_cam,
                                          const double *edge_uv, const double K_left[4], const double K_right[4],
                                          const double ext_left[7], const double ext_right[7], double huber_delta,
                                          int jacobian_mode);
SVS_API void svs_ba_shard_destroy(svs_ctx *ctx, svs_ba_shard *sh);
SVS_API int svs_ba_shard_window(const svs_ba_shard *sh, void **window_dev, size_t *bytes);
/* peer_window_dev[r] = rank r's window as a device pointer valid in THIS process; peer_window_dev[rank] must be the own one */
SVS_API int svs_ba_shard_set_peers(svs_ctx *ctx, svs_ba_shard *sh, int n_ranks, int rank, void *const *peer_window_dev);
/* several shards on ONE GPU (tests): bound each kernel's grid so that all of them are resident at the same time */
SVS_API int svs_ba_shard_set_grid_limit(svs_ba_shard *sh, int max_ctas);
SVS_API int svs_ba_shard_optimize(svs_ctx *ctx, svs_ba_shard *sh, int max_iter, svs_ba_stats *stats);
SVS_API int svs_ba_shard_launch(svs_ctx *ctx, svs_ba_shard *sh, int max_iter);        /* asynchronous half of _optimize */
SVS_API int svs_ba_shard_finish(svs_ctx *ctx, svs_ba_shard *sh, svs_ba_stats *stats); /* waits, returns the statistics */
SVS_API int svs_ba_shard_get(svs_ctx *ctx, svs_ba_shard *sh, double *poses_out, double *lms_out, double *edge_chi2_out);
/* diagnostics: nanoseconds per solver phase of the last optimize, measured in the kernel (phase list in csrc/ba_shard.cu) */
SVS_API int svs_ba_shard_phase_ns(svs_ctx *ctx, svs_ba_shard *sh, double out[16]);
/* CUDA IPC plumbing for the windows (cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle): 64-byte handles */
SVS_API int svs_ipc_export(svs_ctx *ctx, const void *dev_ptr, unsigned char handle_out[64]);
SVS_API int svs_ipc_import(svs_ctx *ctx, const unsigned char handle[64], void **dev_ptr_out);
SVS_API int svs_ipc_release(svs_ctx *ctx, void *dev_ptr);

/* ---------------------------------------------------------------- f4 : pose-graph optimisation
 * Replaces the g2o block of LoopClosure::PoseGraphOptimization (src/loopclosure.cpp:641-746): one VertexPose per keyframe,
 * EdgePoseGraph (include/StereoVisionSLAM/g2o_types.h:231-267, error = log(measurement^-1 * pose_a * pose_b^-1), information
 * I6) for every consecutive-keyframe pair (measurement = Frame::relative_pose_pkf_) and every closed loop
 * (Frame::loop_relative_pose_), fixed[k] != 0 for the vertices held constant (keyframe 0 at :693-696), Levenberg-Marquardt
 * over a dense pivoted LDLT of the whole system, optimizer.optimize(max_iter = 22).  Edge e connects vertex edge_a[e]
 * (the later keyframe) to edge_b[e].  jacobian_mode: 1 = g2o's numeric central differences with delta 1e-9 (what the
 * reference does: the edge has no linearizeOplus), 0 = closed form.  The whole loop runs in one cooperative kernel.
 * svs_pose_graph_move_landmarks is the landmark update that follows (:749-777): pos <- new_pose[k]^-1 * (old_pose[k] * pos)
 * with k = lm_kf[i] the keyframe of the landmark's first valid observation (< 0: landmark untouched). */
SVS_API int svs_pose_graph_optimize(svs_ctx *ctx, int n_kf, double *poses_inout /* 7*n_kf */, const uint8_t *fixed /* n_kf */,
                                    int n_edge, const int32_t *edge_a, const int32_t *edge_b, const double *measurement /* 7*n_edge */,
                                    int max_iter, int jacobian_mode, svs_ba_stats *stats);
SVS_API int svs_pose_graph_move_landmarks(svs_ctx *ctx, int n_lm, double *lms_inout /* 3*n_lm */, const int32_t *lm_kf,
                                          int n_kf, const double *old_poses, const double *new_poses);

/* ---------------------------------------------------------------- a10 : dense stereo
 * svs_stereo_bm replaces stereo_depth_est_->compute(l, r, disp) (src/dense_reconstruction.cpp:114) for
 * cv::StereoBM::create(ndisp, block) with OpenCV's defaults (XSOBEL prefilter cap 31, texture 10,
 * uniqueness 15, min disparity 0, no speckle filter): int16 disparity*16, invalid = -16.
 * svs_backproject replaces src/dense_reconstruction.cpp:116-173: disp/16 -> depth = fx*baseline/d (f32),
 * every pixel with depth >= 1 back-projected with Camera::pixel2world (src/camera.cpp:82-86), x-outer /
 * y-inner order, colour from the BGR image. */
SVS_API int svs_stereo_bm(svs_ctx *ctx, const uint8_t *left, const uint8_t *right, int w, int h, int stride,
                          int n_images, size_t img_stride_bytes, int ndisp, int block,
                          int16_t *disp_out /* n*h*w */);
/* Same with the images and the disparity maps resident in device memory (a batch already in HBM, e.g. straight from a frame
 * set or a decoder); synchronises the context stream before returning. */
SVS_API int svs_stereo_bm_dev(svs_ctx *ctx, const uint8_t *left_dev, const uint8_t *right_dev, int w, int h, int stride,
                              int n_images, size_t img_stride_bytes, int ndisp, int block, int16_t *disp_out_dev);
SVS_API int svs_backproject(svs_ctx *ctx, const int16_t *disp, const uint8_t *bgr /* h*w*3 */, int w, int h,
                            const double K[4], double baseline, const double cam_pose_inv[7],
                            const double T_cw[7], float *xyz_out /* 3*h*w */, uint8_t *rgb_out /* 3*h*w */,
                            int32_t *n_out);
SVS_API int svs_bgr2gray(svs_ctx *ctx, const uint8_t *bgr, int w, int h, int n_images, uint8_t *gray);

/* ---------------------------------------------------------------- f3 : dense-map post-processing
 * svs_pointcloud_sor replaces pcl::StatisticalOutlierRemoval with setMeanK(mean_k = 50), setStddevMulThresh(stddev_mul = 1.0) as
 * used per keyframe and on the merged map (src/dense_reconstruction.cpp:179-184, :194-200): the mean distance of every point to
 * its mean_k nearest neighbours (exact k-NN, float distances), keep[i] = mean_i <= mu + stddev_mul * sigma over all points.
 * svs_voxel_grid replaces pcl::VoxelGrid with setLeafSize(leaf = 0.02) (:203-209): one centroid per occupied voxel (xyz and
 * colour averaged), in ascending voxel index; like PCL, a cloud whose voxel index would overflow 32 bits is returned
 * unchanged (n_out = n).  PCL is un-vendored: semantics restated from PCL 1.12, parity unpinned. */
SVS_API int svs_pointcloud_sor(svs_ctx *ctx, const float *xyz /* 3n */, int n, int mean_k, double stddev_mul, uint8_t *keep_out /* n */,
                               float *mean_dist_out /* n or NULL */, int *n_kept);
SVS_API int svs_voxel_grid(svs_ctx *ctx, const float *xyz /* 3n */, const uint8_t *rgb /* 3n or NULL */, int n, double leaf,
                           float *xyz_out /* 3n */, uint8_t *rgb_out /* 3n or NULL */, int *n_out);

/* ---------------------------------------------------------------- a6 / a8 / a9 : the pipeline
 * svs_slam steps n_streams independent stereo streams in lock-step through the host-side mirror of the reference's
 * Frontend / Backend / Map classes (stereovision-slam_b200/host/slam.h), i.e. it is Frontend::AddFrame
 * (src/frontend.cpp:690-721) for a batch of streams: every third-party seam becomes one batched svs_* call per step.
 * Bundle adjustment runs on the synchronous schedule (inside Backend::UpdateMap).  Config fields mirror
 * config/stereo_slam_configs/default.yaml plus the constants hard-coded at src/frontend.cpp:24,107-108 and
 * src/backend.cpp:164. */
typedef struct {
    int32_t num_features, num_features_init, num_features_tracking, num_features_tracking_bad;
    int32_t num_features_needed_for_keyframe, num_active_keyframes, backend_on;
    int32_t lk_win, lk_max_level, lk_max_iter, ba_max_iter, ba_jacobian_mode, oracle_simd_granule;
    double max_triangulation_depth, chi2_th, gftt_quality, gftt_min_distance, lk_eps;
    /* 1 (default): right images are ingested only for the streams that insert a keyframe in this step (identical results,
     * ~half the image traffic); 0: both eyes of every frame are ingested, as Dataset::NextFrame resizes both */
    int32_t lazy_right_ingest;
    /* 1 (default): Frontend::Track()'s per-frame arithmetic (motion model, LK initial guesses, LK, feature hand-over, pose-only LM,
     * status / keyframe test, relative motion) runs on device-resident per-stream state with no host round trip between the
     * seams; the host classes see a stream only when it inserts a keyframe or initialises.  0: every seam is a host round trip
     * (bit-identical results) */
    int32_t device_tracking;
} svs_slam_config;
typedef struct svs_slam svs_slam;
SVS_API void svs_slam_default_config(svs_slam_config *cfg);
/* K = intrinsics of the PROCESSED image (already halved when half != 0, src/dataset.cpp:73); right camera extrinsic
 * is the pure translation (-baseline, 0, 0) (src/dataset.cpp:63-77). */
SVS_API svs_slam *svs_slam_create(svs_ctx *ctx, int n_streams, int in_w, int in_h, int half, const svs_slam_config *cfg,
                                  const double K[4], double baseline);
SVS_API void svs_slam_destroy(svs_slam *s);
/* One Frontend::AddFrame for every stream.  status: 0 INITING, 1 TRACKING_GOOD, 2 TRACKING_BAD, 3 LOST. */
SVS_API int svs_slam_add_frames(svs_slam *s, const uint8_t *const *left, const uint8_t *const *right, size_t row_stride,
                                int on_device, double *poses_out /* 7*n */, int32_t *status_out, int32_t *keyframe_out,
                                int32_t *inliers_out);
/* Optional hint: the frames of the NEXT svs_slam_add_frames call after the coming one (n pointers each, same row stride
 * and on_device as the coming call).  The coming call starts their ingest (svs_frameset_prefetch_ptrs) right after its
 * own push, overlapping it with this step's kernels.  Consumed by one call. */
SVS_API int svs_slam_hint_next(svs_slam *s, const uint8_t *const *next_left, const uint8_t *const *next_right);
SVS_API int svs_slam_get_features(svs_slam *s, int stream, int right, float *xy, int64_t *map_point_ids, uint8_t *valid,
                                  int cap, int *n);
SVS_API int svs_slam_get_keyframes(svs_slam *s, int stream, int active_only, int64_t *kf_ids, int64_t *frame_ids,
                                   double *poses, int cap, int *n);
SVS_API int svs_slam_get_landmarks(svs_slam *s, int stream, int active_only, int64_t *ids, double *xyz,
                                   int32_t *observed_times, int cap, int *n);
/* phase_seconds[8]: push, track-LK, pose LM, detect, right-LK, triangulate, BA, host bookkeeping (wall clock, includes
 * device time).  counters[12]: frames, keyframes, BA problems, BA iterations, BA trials, BA edges, BA landmarks,
 * BA keyframes, LK points, pose-LM edges, image bytes read from host memory, right images ingested (12 values). */
SVS_API int svs_slam_get_counters(svs_slam *s, double *phase_seconds, long long *counters);
/* host_seconds[8]: the "host bookkeeping" phase split by section (begin + prepare track, finish track + prepare pose,
 * finish pose + prepare detect, finish detect + prepare right, finish right + prepare triangulate, finish triangulate +
 * prepare BA, finish BA + end, unused). */
SVS_API int svs_slam_get_host_seconds(svs_slam *s, double *host_seconds);
SVS_API svs_frameset *svs_slam_frameset(svs_slam *s);
/* Host threads (OpenMP) this pipeline uses for its per-stream bookkeeping; several pipelines on distinct contexts may be
 * stepped concurrently from different host threads (their kernels and copies overlap on the device). */
SVS_API int svs_slam_set_threads(svs_slam *s, int n);

/* ---------------------------------------------------------------- data formats either side of the path (host only)
 * svs_kitti_read_calib     Dataset::initialize (src/dataset.cpp:24-80): calib.txt -> per camera (fx fy cx cy), translation
 *                          t = K^-1 P[:,3] and baseline |t|; K is halved when half != 0 (the reference always does, :73).
 * svs_write_keyframes_txt  / svs_write_landmarks_pcd: the two files of VisualOdometry::saveSLAMOutputInFile
 *                          (src/visual_odometry.cpp:198-310), from the arrays svs_slam_get_keyframes / _landmarks return. */
SVS_API int svs_kitti_read_calib(const char *calib_path, int half, double K_out[16], double t_out[12], double baseline_out[4]);
SVS_API int svs_write_keyframes_txt(const char *path, const char *dataset_dir, int left_cam_index, int n, const int64_t *frame_ids,
                                    const double *poses /* 7n */);
SVS_API int svs_write_landmarks_pcd(const char *path, int n, const double *xyz /* 3n */);

#ifdef __cplusplus
}
#endif
#endif /* SVSLAM_H */
