/*
 * ssm.h -- C ABI of libssm.so: the B200-native (sm_100a) dense stereo-to-semantic-map path.
 *
 * Drop-in boundary for the reference's two entry points on this path:
 *   calDisparity_SGBM(img_L, img_R, disp)      /root/reference include/stereo.h:15, src/stereo.cpp:11-38
 *   rgbd_tutor::Mapper (generatePointCloud,    /root/reference include/mapper.h:15-70,
 *     semantic_motion_fuse, viewer's VoxelGrid   src/mapper.cpp:12-94,96-178,189-216
 *     fusion)
 * plus the glue between them (FrameReader::next disparity->depth, src/rgbdframe.cpp:85-116, and
 * RGBDFrame::project2dTo3d, include/rgbdframe.h:63-75).
 *
 * Conventions: extern "C"; plain pointers and sizes; every call returns an ssm_status code: 0 = ok,
 * <0 = error with the message in ssm_last_error; no exceptions cross the boundary; no OpenCV / PCL /
 * Eigen / torch types.  "host" entry points take host pointers and block; "device" entry points
 * take device pointers on the ctx's GPU and are asynchronous on the given cudaStream_t (passed
 * as void*; NULL = the ctx's own stream).  One ssm_ctx per GPU; calls on a ctx are serialised
 * by the caller (the reference runs SGBM on its main thread and mapping on the viewer thread:
 * use two contexts or a lock).  There is no CPU fallback: ssm_create fails when no sm_100
 * device is present.
 *
 * Image layouts match cv::Mat: row-major, `stride` in BYTES between rows; 8UC3 is interleaved BGR.
 */
#ifndef SSM_H
#define SSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSM_MAX_LABELS 20 /* 12 (SegNet) and 19 (Cityscapes) fit one 128-byte voxel record */
#define SSM_LABEL_UNKNOWN 255

typedef enum {
    SSM_OK = 0,
    SSM_ERR_INVALID_ARGUMENT = -1, /* bad shape / parameter (the reference would throw cv::Exception) */
    SSM_ERR_CUDA = -2,             /* CUDA runtime error; text in ssm_last_error */
    SSM_ERR_NO_DEVICE = -3,        /* no sm_100 GPU: there is no CPU fallback */
    SSM_ERR_CAPACITY = -4,         /* frame larger than ctx limits, or voxel hash full with no memory left to grow */
    SSM_ERR_COMM = -5,             /* NCCL error / communicator not initialised */
    SSM_ERR_UNSUPPORTED = -6       /* parameter combination outside the 16-bit cost range */
} ssm_status;

/* All tunables of the path in one POD (SURVEY.md section 5 "Config / flags"). */
typedef struct ssm_params {
    /* cv::StereoSGBM fields as set by src/stereo.cpp:16-28 */
    int min_disparity;       /* 0 (only 0 is supported) */
    int num_disparities;     /* 80 in the reference; 128 / 256 in BASELINE.json configs; multiple of 16 */
    int block_size;          /* SADWindowSize = 11 */
    int p1, p2;              /* 4*11*11, 32*11*11 */
    int disp12_max_diff;     /* 1 */
    int pre_filter_cap;      /* 63 */
    int uniqueness_ratio;    /* 10 */
    int speckle_window_size; /* 100 */
    int speckle_range;       /* 32 */
    /* camera: parameters.txt:37-41,50-54,63 (read at src/rgbdframe.cpp:87-94) */
    double cx, cy, fx, fy, baseline, scale;
    double roix, roiy, roiz;
    /* mapper: parameters.txt:97-98 (include/mapper.h:24-25) */
    double resolution;   /* voxel leaf, metres */
    double max_distance; /* metres */
    /* labels: class id -> semantic BGR colour (src/mapper.cpp:42-54) */
    int num_labels;
    uint8_t palette_bgr[SSM_MAX_LABELS][3];
    uint32_t drop_mask;      /* bit l set: class l is removed from the cloud (mapper.cpp:41-55) */
    uint32_t dynamic_mask;   /* bit l set: class l seeds the dilated moving mask (mapper.cpp:206-208) */
    int dilate_iterations;   /* mapper.cpp:214 */
    int colour_source;       /* 0: left rgb image (mapper.cpp:72-84); 1: semantic colour (mapper.cpp~:60) */
    /* capacities (allocation limits of the context) */
    int max_width, max_height, max_batch;
    uint64_t map_capacity;   /* initial voxel hash slots (rounded up to a power of two); the table doubles when half full */
} ssm_params;

typedef struct ssm_ctx ssm_ctx;

/* One fused voxel, as exported.  xyz = centroid, rgba = 0x00RRGGBB of the truncated mean colour
 * (pcl::VoxelGrid semantics), label = majority vote (ties -> lowest id). */
typedef struct ssm_voxel_export {
    int32_t* ijk;     /* [n][3] voxel coordinates floor(coord / leaf) */
    float* xyz;       /* [n][3] */
    uint32_t* rgba;   /* [n]    */
    uint8_t* label;   /* [n]    */
    uint32_t* count;  /* [n]    */
    uint32_t* votes;  /* [n][num_labels] */
} ssm_voxel_export;   /* any member may be NULL */

/* ---- lifecycle ------------------------------------------------------------------------------ */
void ssm_default_params(ssm_params* p); /* reference defaults (stereo.cpp:16-28, parameters.txt) */
int ssm_create(const ssm_params* p, int device, ssm_ctx** out);
void ssm_destroy(ssm_ctx* ctx);
const char* ssm_last_error(void);
const char* ssm_version(void);
/* number of kernels this library launched on the ctx since creation (for bench.py gpu_launches) */
uint64_t ssm_kernel_launches(const ssm_ctx* ctx);
/* per-stage device time of the most recent *_host / pipeline call, ms (stage ids below) */
enum { SSM_STAGE_COST = 0, SSM_STAGE_VERTICAL = 1, SSM_STAGE_HORIZONTAL = 2, SSM_STAGE_SELECT = 3, SSM_STAGE_POST = 4,
       SSM_STAGE_POINTS = 5, SSM_STAGE_FUSE = 6, SSM_STAGE_COUNT = 7 };
int ssm_set_stage_timing(ssm_ctx* ctx, int enabled);
int ssm_stage_time_ms(ssm_ctx* ctx, int stage, float* ms);

/* ---- stereo.h: calDisparity_SGBM -------------------------------------------------------------- */
/* Host, blocking.  left/right 8UC1 w x h; disp 16SC1 w x h (disparity x16, invalid = -16). */
int ssm_sgbm(ssm_ctx* ctx, const uint8_t* left, const uint8_t* right, int w, int h, size_t stride,
             int16_t* disp, size_t disp_stride);
/* Device, async.  Densely packed [batch][h][w] buffers. */
int ssm_sgbm_batch_device(ssm_ctx* ctx, int batch, const uint8_t* d_left, const uint8_t* d_right, int w, int h,
                          int16_t* d_disp, void* stream);
/* Debug/parity taps (device buffers, valid after ssm_sgbm*, batch item 0..): matching cost C and the
 * aggregated cost S_f = sat(L0+L1+L2+L3) (the fifth path is added in registers only), both
 * [batch][h][w-D][D] int16; raw WTA disparity after the L-R check; disparity after the median. */
int ssm_debug_copy_volume(ssm_ctx* ctx, int which /*0=C,1=S,2=disp_raw,3=disp_median*/, int batch_index,
                          void* host_dst, size_t bytes);

/* ---- FrameReader::next glue: disparity -> depth (rgbdframe.cpp:85-116) ----------------------- */
int ssm_disparity_to_depth(ssm_ctx* ctx, const int16_t* disp, int w, int h, size_t disp_stride,
                           uint16_t* depth, size_t depth_stride);

/* ---- mapper.h ----------------------------------------------------------------------------------- */
/* Mapper::semantic_motion_fuse: semantic 8UC3 BGR -> moving mask 8UC1 (255 = moving). */
int ssm_semantic_motion_fuse(ssm_ctx* ctx, const uint8_t* semantic_bgr, int w, int h, size_t stride,
                             uint8_t* mask, size_t mask_stride);
/* Mapper::generatePointCloud: depth 16UC1 + semantic/rgb 8UC3 + pose (row-major 4x4 double,
 * camera->world = frame->T_f_w) -> row-major-ordered cloud.  Outputs hold up to max_points; the
 * number produced is returned in *n_points.  xyz [n][3] fp32 world frame; rgba 0x00RRGGBB; label id. */
int ssm_generate_point_cloud(ssm_ctx* ctx, const uint16_t* depth, const uint8_t* semantic_bgr,
                             const uint8_t* rgb_bgr, int w, int h, const double* T_f_w,
                             float* xyz, uint32_t* rgba, uint8_t* label, int max_points, int* n_points);
/* Mapper::viewer accumulate + VoxelGrid: fuse one frame (host buffers) into the global map. */
int ssm_map_integrate_frame(ssm_ctx* ctx, const uint16_t* depth, const uint8_t* semantic_bgr,
                            const uint8_t* rgb_bgr, int w, int h, const double* T_f_w);
/* fuse an explicit cloud (e.g. one returned by ssm_generate_point_cloud) */
int ssm_map_integrate_points(ssm_ctx* ctx, const float* xyz, const uint32_t* rgba, const uint8_t* label, int n);
/* Keyframe cache + redraw.  Mapper::generatePointCloud caches each keyframe's cloud in camera coordinates
 * (frame->pointcloud, mapper.cpp:17-20) and re-transforms it by the frame's CURRENT pose on every update (:90-91),
 * because the pose graph rewrites keyframe poses (pose_graph.cpp:253-260).  ssm_keyframe_add builds and keeps that
 * cloud on the device and returns its id; ssm_keyframe_set_pose records a new T_f_w; ssm_map_redraw is the periodic
 * full redraw (mapper.cpp:121-131: globalMap->clear(), then += the listed keyframes -- the reference passes every
 * second one) and ssm_map_integrate_keyframes the incremental branch (:132-149).  ids == NULL: every live keyframe. */
int ssm_keyframe_add(ssm_ctx* ctx, const uint16_t* depth, const uint8_t* semantic_bgr, const uint8_t* rgb_bgr,
                     int w, int h, const double* T_f_w, int* id_out);
int ssm_keyframe_set_pose(ssm_ctx* ctx, int id, const double* T_f_w);
int ssm_keyframe_release(ssm_ctx* ctx, int id);
int ssm_keyframe_count(ssm_ctx* ctx, int* n_live, uint64_t* n_points);
int ssm_map_redraw(ssm_ctx* ctx, const int* ids, int n);
int ssm_map_integrate_keyframes(ssm_ctx* ctx, const int* ids, int n);
int ssm_map_clear(ssm_ctx* ctx);                      /* globalMap->clear(), mapper.cpp:125 */
int ssm_map_size(ssm_ctx* ctx, uint64_t* n_voxels);   /* "points in global map", mapper.cpp:161 */
/* Export up to max_voxels voxels of THIS rank's table; sorted != 0 orders them by (k,j,i)
 * (pcl::VoxelGrid output order).  *n_out receives the number of voxels in the map.  Everything runs on the device -- K9
 * voxel_finalize (centroid division, truncated mean colour, majority label) and a radix sort into PCL's order -- followed
 * by one D2H copy per requested array (pinned destinations make those asynchronous DMA). */
int ssm_map_export(ssm_ctx* ctx, const ssm_voxel_export* out, uint64_t max_voxels, int sorted, uint64_t* n_out);
/* Device time (ms) of the kernels of the latest ssm_map_export* call on this context: K9 voxel_finalize + the radix sort. */
int ssm_map_export_device_ms(ssm_ctx* ctx, float* ms);
/* Multi-GPU export (SURVEY 8e "Export"): collective over the communicator -- every rank calls it with the same `sorted`;
 * the ranks' tables (disjoint by ownership) travel to rank 0 over NCCL, which finalizes and orders the union and fills
 * `out` (other ranks may pass out == NULL and receive *n_out = 0).  Without a communicator == ssm_map_export. */
int ssm_map_export_gathered(ssm_ctx* ctx, const ssm_voxel_export* out, uint64_t max_voxels, int sorted, uint64_t* n_out);
/* The voxel hash grows by itself (doubling, stream-ordered, no host stall; DESIGN.md "growth"); ssm_map_reserve grows it
 * ahead of time to at least `slots` slots.  ssm_map_stats reports its state (blocking). */
typedef struct ssm_map_statistics {
    uint64_t slots;        /* table size */
    uint64_t voxels;       /* occupied slots */
    double load_factor;    /* voxels / slots */
    double mean_probe;     /* mean displacement of a record from its home slot (linear probing) */
    uint64_t max_probe;    /* longest displacement */
    uint64_t grow_steps;   /* growth steps since ssm_create */
    uint64_t table_bytes;  /* 128 bytes per slot */
} ssm_map_statistics;
int ssm_map_reserve(ssm_ctx* ctx, uint64_t slots);
int ssm_map_stats(ssm_ctx* ctx, ssm_map_statistics* out);
/* Write the fused map as a binary PCD with fields x y z rgba (pcl::PCDWriter, mapper.cpp:165-170). */
int ssm_map_save_pcd(ssm_ctx* ctx, const char* path);

/* ---- dense motion cues: the other dense consumer of the disparity map (tracker thread, src/track.cpp:67-79) ---- */
/* The record image `xyz` is [h][w][10] fp32 = X, Y, Z, u, v, disparity, intensity, I_u, I_v, motion mark (CV_32FC(10)).
 * Host entry points take host pointers and block; they mirror the reference functions one to one. */
/* triangulate10D(img, disp, xyz, f, cx, cy, b, roi): src/stereo.cpp:41-118 (include/stereo.h:25) */
int ssm_triangulate10d(ssm_ctx* ctx, const uint8_t* img, size_t img_stride, const int16_t* disp, size_t disp_stride,
                       int w, int h, double f, double cx, double cy, double b, double roi_x, double roi_y, double roi_z,
                       float* xyz);
/* correct3DPoints(xyz, roi, pitch1, pitch2): src/stereo.cpp:127-181 (include/stereo.h:36); in place */
int ssm_correct_3d_points(ssm_ctx* ctx, float* xyz, int w, int h, double roi_x, double roi_y, double roi_z,
                          double pitch1, double pitch2);
/* setImageROI(xyz, roi_mask): src/stereo.cpp:183-192 (include/stereo.h:43) */
int ssm_set_image_roi(ssm_ctx* ctx, const float* xyz, int w, int h, uint8_t* roi_mask, size_t mask_stride);
/* UVDisparity::calVDisparity(img_dis, xyz): src/uvdisparity.cpp:277-366 (include/uvdisparity.hpp:91).  *v_cols =
 * cvCeil(max(disp)/16); v_dis_int [h][v_cols] int32 and v_dis [h][v_cols] u8 are dense (any may be NULL; both need
 * v_cols <= cap_cols); channel 8 of xyz (may be NULL) is filled. */
int ssm_v_disparity(ssm_ctx* ctx, const int16_t* disp, size_t disp_stride, int w, int h, float* xyz,
                    int32_t* v_dis_int, uint8_t* v_dis, int cap_cols, int* v_cols);
/* UVDisparity::calUDisparity(img_dis, xyz, roi_mask, ground_mask): src/uvdisparity.cpp:195-274
 * (include/uvdisparity.hpp:88).  *u_rows = cvCeil(max(disp)/16) + 1; u_dis_int / u_dis [u_rows][w]; channel 7 of xyz. */
int ssm_u_disparity(ssm_ctx* ctx, const int16_t* disp, size_t disp_stride, int w, int h, float* xyz,
                    const uint8_t* roi_mask, const uint8_t* ground_mask, int32_t* u_dis_int, uint8_t* u_dis,
                    int cap_rows, int* u_rows);
/* The same chain on device-resident batches, in the order of UVDisparity::Process (src/uvdisparity.cpp:842-903), two
 * asynchronous calls around the host's pitch estimation:
 *   stage 1 = triangulate10D + calVDisparity: xyz [batch][h][w][10] complete up to channel 8; optional per-frame V maps
 *             at d_v_dis_int / d_v_dis + frame * hist_stride, frame-local layout [h][v_cols] (v_cols <= cap_cols,
 *             hist_stride >= h * cap_cols + 4 elements);
 *   stage 2 = correct3DPoints + setImageROI + calUDisparity: xyz updated in place (channels 1, 2, 6, 7), roi mask
 *             [batch][h][w], optional U maps [u_rows][w] per frame (u_rows <= cap_rows, hist_stride >= cap_rows * w).
 *             d_ground_mask NULL = all ground.
 * ssm_motion_cues_overflow reports (after synchronisation) bit 0: a V map exceeded cap_cols, bit 1: a U map cap_rows. */
int ssm_motion_cues_stage1_device(ssm_ctx* ctx, int batch, const uint8_t* d_img, const int16_t* d_disp, int w, int h,
                                  double f, double cx, double cy, double b, float* d_xyz, int32_t* d_v_dis_int,
                                  uint8_t* d_v_dis, size_t hist_stride, int cap_cols, void* stream);
int ssm_motion_cues_stage2_device(ssm_ctx* ctx, int batch, const int16_t* d_disp, int w, int h, float* d_xyz,
                                  double roi_x, double roi_y, double roi_z, double pitch1, const uint8_t* d_ground_mask,
                                  uint8_t* d_roi_mask, int32_t* d_u_dis_int, uint8_t* d_u_dis, size_t hist_stride,
                                  int cap_rows, void* stream);
int ssm_motion_cues_overflow(ssm_ctx* ctx, int* flags);

/* ---- label production in front of the path (experiment/segnet.cpp:121-135, src/rgbdframe.cpp:118-136) -------- */
/* SegNet's argmax index image (8UC1, sw x sh; 480 x 360 in the reference) -> cv::resize to the frame size (default
 * INTER_LINEAR) -> cv::LUT through the 256-entry BGR colour table -> the semantic image the Mapper reads.  `raw`
 * (optional) receives the resized index image (= frame->raw_semantic).  Bit-exact with cv2.resize / cv2.LUT. */
int ssm_labels_from_indices(ssm_ctx* ctx, const uint8_t* index, size_t index_stride, int sw, int sh, int dw, int dh,
                            const uint8_t* lut_bgr /* [256][3] */, uint8_t* semantic_bgr, size_t sem_stride,
                            uint8_t* raw, size_t raw_stride);
/* Device, async: d_index [batch][sh][sw] -> d_semantic_bgr [batch][dh][dw][3] (+ d_raw [batch][dh][dw], may be NULL);
 * the output feeds ssm_pipeline_batch_device directly.  lut_bgr is a host pointer. */
int ssm_labels_from_indices_batch_device(ssm_ctx* ctx, int batch, const uint8_t* d_index, int sw, int sh, int dw, int dh,
                                         const uint8_t* lut_bgr, uint8_t* d_semantic_bgr, uint8_t* d_raw, void* stream);

/* ---- PNG ingest in front of the path (FrameReader::next, src/rgbdframe.cpp:45-78, 138-180) ----- */
/* The reference reads every image with cv::imread: the grey stereo pair with flag 0, the colour and label images with
 * the default flag.  ssm_png_decode_batch_device decodes `batch` PNG files held in host memory into device images of
 * the layout the pipeline entry points take: mode 0 -> [batch][h][w] u8 (== imread(path, 0)), mode 1 -> [batch][h][w][3]
 * u8 BGR (== imread(path)).  host_threads = 0: the zlib streams are inflated ON THE GPU, one warp per file (the host threads only
 * walk the chunks, verify their CRCs and gather the IDAT payloads; SSM_HOST_INFLATE=1 in the environment turns this off);
 * host_threads >= 1: they are inflated with zlib on that many host threads.  PNG
 * un-filtering, palette expansion, alpha stripping, channel reordering and the colour -> grey conversion run on the GPU
 * on `stream` (NULL = the context's stream).  Every file must be w x h, 8 bits per sample, non-interlaced (grey,
 * grey + alpha, RGB, RGBA or palette).  Bit-exact with cv2 4.13.  The host returns when the batch is queued.  Container
 * errors (signature, chunk CRC, size, colour type) are returned by the call itself; with the GPU decoder a corrupt or short
 * zlib stream or a scanline filter type above 4 is only known when the batch has run: it is returned by ssm_png_batch_wait,
 * or by the fourth ssm_png_decode_batch_device call after it (the one that re-uses the batch's staging buffers). */
int ssm_png_info(const uint8_t* png, size_t png_bytes, int* w, int* h, int* channels /* 1 or 3, may be NULL */);
int ssm_png_decode_batch_device(ssm_ctx* ctx, int batch, const uint8_t* const* png, const size_t* png_bytes, int w, int h,
                                int mode, uint8_t* d_out, int host_threads, void* stream);
/* Blocks until every queued ingest batch has been decoded; returns the first stream error of a GPU-inflated batch, if any. */
int ssm_png_batch_wait(ssm_ctx* ctx);
/* The GPU decoder on its own (blocking; test and tooling entry point): n zlib streams (RFC 1950 / 1951) in host memory ->
 * out[i][0 .. out_bytes[i]) in host memory.  status[i]: 0 = ok, 1 bad zlib header, 2 bad block, 3 bad code lengths, 4 bad
 * symbol, 5 distance too far back, 6 input ends early, 7 Adler-32 mismatch, 8 fewer than out_bytes[i] bytes in the stream.
 * As with inflate() and a full output buffer, data beyond out_bytes[i] is ignored. */
int ssm_zlib_inflate_batch(ssm_ctx* ctx, int n, const uint8_t* const* streams, const size_t* stream_bytes, uint8_t* const* out,
                           const size_t* out_bytes, int* status);
/* one file, host buffer out (blocking): == cv::imread(path, mode ? 1 : 0) */
int ssm_png_decode(ssm_ctx* ctx, const uint8_t* png, size_t png_bytes, int mode, uint8_t* out, size_t out_bytes, int* w, int* h);

/* ---- the whole path, batched (north_star: stereo pair + labels + pose -> map) ----------------- */
/* Device-resident inputs: [batch][h][w] u8 left/right, [batch][h][w][3] u8 semantic/rgb BGR,
 * [batch][16] double poses.  d_disp_out (optional, may be NULL) receives [batch][h][w] int16. */
int ssm_pipeline_batch_device(ssm_ctx* ctx, int batch, const uint8_t* d_left, const uint8_t* d_right,
                              const uint8_t* d_semantic, const uint8_t* d_rgb, const double* d_poses,
                              int w, int h, int16_t* d_disp_out, void* stream);
/* Same through host buffers (pinned or pageable): H2D copies, the path, and a D2H of the map size. */
int ssm_pipeline_batch_host(ssm_ctx* ctx, int batch, const uint8_t* left, const uint8_t* right,
                            const uint8_t* semantic, const uint8_t* rgb, const double* poses,
                            int w, int h, int16_t* disp_out /* may be NULL */, uint64_t* n_voxels_out);
/* Asynchronous variant for streaming callers: returns as soon as the copies and kernels are enqueued.  Inputs are
 * staged through two device buffer sets on a copy stream, so the H2D copy of call k+1 overlaps the kernels of call
 * k.  The host buffers must stay valid (and should be pinned) until ssm_synchronize returns or two further calls have
 * been made.  n_voxels_pinned (optional, pinned host memory) receives the map size after this batch, by an
 * asynchronous D2H copy ordered after the batch's kernels.  Errors of the voxel hash (capacity) surface at the next
 * ssm_map_size / ssm_map_export / blocking call.  The pipeline entry points let the host run at most two batches ahead
 * of the device (a call waits for the batch three calls back): that is what lets the library size the voxel hash from
 * the counts of completed batches before the table fills. */
int ssm_pipeline_batch_host_async(ssm_ctx* ctx, int batch, const uint8_t* left, const uint8_t* right,
                                  const uint8_t* semantic, const uint8_t* rgb, const double* poses, int w, int h,
                                  uint32_t* n_voxels_pinned);
int ssm_synchronize(ssm_ctx* ctx);

/* ---- multi-GPU: spatially owned voxel hash, NCCL all-to-all point routing ---------------------- */
#define SSM_UNIQUE_ID_BYTES 128
int ssm_comm_get_unique_id(uint8_t id[SSM_UNIQUE_ID_BYTES]);               /* rank 0, then broadcast by the host */
int ssm_comm_init(ssm_ctx* ctx, const uint8_t id[SSM_UNIQUE_ID_BYTES], int rank, int nranks);
/* Peer-memory routing (preferred on one NVLink/NVSwitch box): after ssm_comm_init every rank exports its inbox as a
 * CUDA IPC handle, the host all-gathers the handles ([nranks][SSM_IPC_HANDLE_BYTES]) and every rank connects.  From
 * then on point generation and the dispatch all-to-all are ONE kernel (one 16-byte peer store over NVLink per routed point),
 * followed by a stream-ordered step barrier over arrival flags in the peers' inbox headers (no collective, no host
 * synchronisation per batch) and the fusion of the rank's inbox.  Every rank must make the same sequence of pipeline /
 * integrate calls; a rank that waits more than ~10 s for a peer's flag gives up and the next blocking map call returns
 * SSM_ERR_COMM (its map then lacks that step's remote points).  Without this call the NCCL send/recv all-to-all is used. */
#define SSM_IPC_HANDLE_BYTES 64
int ssm_comm_ipc_export(ssm_ctx* ctx, uint8_t handle[SSM_IPC_HANDLE_BYTES]);
int ssm_comm_ipc_connect(ssm_ctx* ctx, const uint8_t* handles, int nranks);
int ssm_comm_destroy(ssm_ctx* ctx);
/* Streaming option for batched calls with a communicator (default off): the per-batch exchange (points to their owning
 * rank, barrier, inbox fusion) runs on an internal stream, so a pipeline call's stream only covers the stereo half and the
 * next batch's SGBM overlaps this batch's exchange.  The batch is in the map after ssm_synchronize (or any ssm_map_*
 * call, which wait for it).  Every rank must use the same setting. */
int ssm_set_route_overlap(ssm_ctx* ctx, int enabled);
/* owner rank of a voxel (pure function of ijk, brick shift and nranks; exposed for host-side tests) */
int ssm_voxel_owner(int32_t i, int32_t j, int32_t k, int nranks);

#ifdef __cplusplus
}
#endif
#endif /* SSM_H */
