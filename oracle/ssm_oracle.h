/*
 * ssm_oracle.h -- CPU ORACLE for the dense stereo-to-semantic-map path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (semantic_slam_mapping_b200/libssm.so) never links, loads or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's algorithm for the hot path:
 *   - src/stereo.cpp:11-38        calDisparity_SGBM -> cv::StereoSGBM (OpenCV, un-vendored
 *                                 third-party dependency; README.md:48 pins 2.4.x; the only
 *                                 executable copy here is cv2 4.13 MODE_SGBM)
 *   - src/rgbdframe.cpp:85-116    disparity -> ushort depth image
 *   - include/rgbdframe.h:63-75   RGBDFrame::project2dTo3d
 *   - src/mapper.cpp:12-94        Mapper::generatePointCloud
 *   - src/mapper.cpp:189-216      Mapper::semantic_motion_fuse
 *   - src/mapper.cpp:96-178       Mapper::viewer -> pcl::VoxelGrid fusion (PCL 1.7, un-vendored;
 *                                 semantics restated from SURVEY.md App. B)
 *
 *   - src/stereo.cpp:41-192      triangulate10D, correct3DPoints, setImageROI
 *   - src/uvdisparity.cpp:195-366 UVDisparity::calUDisparity / calVDisparity
 *   (label production, experiment/segnet.cpp:121-135, is restated in numpy in oracle/__init__.py)
 *
 * Parity pinning: the SGBM chain is pinned against cv2 4.13 (tests/test_oracle_golden.py and
 * the committed vectors in tests/golden/); triangulate10D / correct3DPoints / setImageROI,
 * calDisparity_SGBM's parameter block and the U/V-disparity histograms against the reference's
 * OWN src/stereo.cpp and src/uvdisparity.cpp, compiled by oracle/Makefile against oracle/cvstub
 * into oracle/_ref/libref_stereo.so (tests/test_oracle_cues.py, tests/golden/cues_ref.npz);
 * cv::resize + cv::LUT against cv2 4.13.  The reference itself ships no tests or golden vectors
 * for this path; PCL cannot be compiled or executed here, so the voxel-fusion part is
 * "parity unpinned" beyond its written definition.
 */
#ifndef SSM_ORACLE_H
#define SSM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int num_disparities;     /* stereo.cpp:18  (80 in the reference; 128/256 in BASELINE configs) */
    int block_size;          /* stereo.cpp:17  SADWindowSize = 11 */
    int p1, p2;              /* stereo.cpp:23-24 */
    int disp12_max_diff;     /* stereo.cpp:28 */
    int pre_filter_cap;      /* stereo.cpp:22 */
    int uniqueness_ratio;    /* stereo.cpp:20 */
    int speckle_window_size; /* stereo.cpp:21 */
    int speckle_range;       /* stereo.cpp:27 */
    int legacy_p2_form;      /* 0: L = C + min(..) - minLr (cv2 4.13, contract); 1: ... - (minLr+P2) (2.4-era) */
} osgbm_params;

typedef struct {
    double cx, cy, fx, fy, baseline, scale; /* parameters.txt:37-41,63 */
    double roix, roiy, roiz;                /* parameters.txt:50-54 */
    double max_distance;                    /* parameters.txt:98 mapper_max_distance */
    int num_labels;                         /* 12 (SegNet) or 19 (Cityscapes) */
    uint8_t palette_bgr[32][3];             /* class id -> semantic BGR colour */
    uint32_t drop_mask;                     /* classes removed from the cloud, mapper.cpp:41-55 */
    uint32_t dynamic_mask;                  /* classes that seed the moving mask, mapper.cpp:206-208 */
    int dilate_iterations;                  /* mapper.cpp:214 (2) */
    int colour_source;                      /* 0 = left rgb image (mapper.cpp:72-84), 1 = semantic colour (mapper.cpp~:60) */
} omap_params;

/* Intermediate volumes are optional (NULL to skip).  C/S are [H][W1][D] int16 with W1 = W - D. */
int oracle_sgbm(const uint8_t* left, const uint8_t* right, int W, int H, size_t stride,
                const osgbm_params* p, int16_t* disp, size_t disp_stride_elems,
                int16_t* C_out, int16_t* S_out, int16_t* disp_raw_out, int16_t* disp_median_out,
                int16_t* Sf_out /* sat(L0+L1+L2+L3), before the right-to-left path is added */,
                int16_t* Sv_out /* sat(L1+L2+L3): the three top-down directions alone */);

void oracle_median3x3_s16(const int16_t* src, int16_t* dst, int W, int H);
void oracle_filter_speckles(int16_t* img, int W, int H, int new_val, int max_speckle_size, int max_diff);

/* rgbdframe.cpp:85-116 */
void oracle_disparity_to_depth(const int16_t* disp, int W, int H, const omap_params* p, uint16_t* depth);

/* mapper.cpp:189-216 with the canonical all-ones 3x3 kernel */
void oracle_moving_mask(const uint8_t* semantic_bgr, int W, int H, const omap_params* p, uint8_t* mask);

/* BGR -> class id via the palette; 255 when the colour is not in the palette */
uint8_t oracle_label_of(const omap_params* p, uint8_t b, uint8_t g, uint8_t r);

/* mapper.cpp:12-94: returns number of points, row-major order.
 * xyz: [n][3] fp32 world frame (after T); xyz_cam: [n][3] camera frame (may be NULL);
 * rgba: packed 0x00RRGGBB | alpha(0xFF<<24 dropped: alpha as stored by PointXYZRGBA default=255);
 * label: class id per point (255 unknown); pix: linear pixel index v*W+u (may be NULL). */
int oracle_generate_point_cloud(const uint16_t* depth, const uint8_t* semantic_bgr, const uint8_t* rgb_bgr,
                                int W, int H, const omap_params* p, const double* T_rowmajor16,
                                float* xyz, float* xyz_cam, uint32_t* rgba, uint8_t* label, int32_t* pix);

/* Voxel map (PCL VoxelGrid semantics + label histogram) */
typedef struct ovoxel_map ovoxel_map;
ovoxel_map* oracle_map_create(double leaf, int num_labels);
void oracle_map_destroy(ovoxel_map* m);
void oracle_map_clear(ovoxel_map* m);
void oracle_map_insert(ovoxel_map* m, const float* xyz, const uint32_t* rgba, const uint8_t* label, int n);
int64_t oracle_map_size(const ovoxel_map* m);
/* export sorted by (k, j, i) ascending == PCL's ascending linear index order.
 * ijk [n][3] int32, centroid [n][3] fp32 (fp32 sequential sums / n, PCL style),
 * centroid_d [n][3] double-accumulated mean, rgba [n] (trunc(mean) packed, alpha 0),
 * count [n], votes [n][num_labels], label [n] (majority, ties -> lowest id, 255 if no votes) */
int64_t oracle_map_export(const ovoxel_map* m, int32_t* ijk, float* centroid, double* centroid_d,
                          uint32_t* rgba, uint32_t* count, uint32_t* votes, uint8_t* label);


/* ---- dense motion cues --------------------------------------------------------- */
/* src/stereo.cpp:41-118; xyz [H][W][10] fp32 */
void oracle_triangulate10d(const uint8_t* img, const int16_t* disp, int W, int H, double f, double cx, double cy, double b,
                           float* xyz);
/* src/stereo.cpp:127-181 (in place) */
void oracle_correct_3d_points(float* xyz, int W, int H, double roi_x, double roi_y, double roi_z, double pitch1, double pitch2);
/* src/stereo.cpp:183-192 */
void oracle_set_image_roi(const float* xyz, int W, int H, uint8_t* roi_mask);
/* src/uvdisparity.cpp:277-366; returns v_cols (or -1 when it exceeds cap_cols); writes xyz channel 8 */
int oracle_v_disparity(const int16_t* disp, int W, int H, float* xyz, int32_t* v_dis_int, uint8_t* v_dis, int cap_cols);
/* src/uvdisparity.cpp:195-274; returns u_rows (or -1 when it exceeds cap_rows); writes xyz channel 7 */
int oracle_u_disparity(const int16_t* disp, int W, int H, float* xyz, const uint8_t* roi_mask, const uint8_t* ground_mask,
                       int32_t* u_dis_int, uint8_t* u_dis, int cap_rows);

#ifdef __cplusplus
}
#endif
#endif
