"""The drop-in sources (semantic_slam_mapping_b200/host/dropin/stereo.cpp, mapper.cpp) are type-checked against the REFERENCE's own
declarations: they include the reference's include/stereo.h and include/mapper.h (from /root/reference) and define exactly what those
headers declare -- calDisparity_SGBM / triangulate10D / correct3DPoints / setImageROI with cv::Mat, and rgbd_tutor::Mapper's
generatePointCloud / semantic_motion_fuse / viewer / SaveMap under the reference's class (constructor
Mapper(const ParameterReader&, PoseGraph&) inline in its header).  OpenCV / PCL / Eigen are not installed in this image, so the
compile runs against oracle/cvstub + oracle/refstub (test-only include path).  Also: host/host.cpp with -DSSM_WITH_OPENCV."""
import os
import subprocess

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SSM_REFERENCE_ROOT", "/root/reference")

needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "include", "mapper.h")), reason="needs the reference's headers")


@needs_reference
def test_dropin_compiles_links_and_defines_the_reference_symbols():
    from semantic_slam_mapping_b200 import build
    build.build()
    path = oracle.build_dropin()
    assert path and os.path.exists(path)
    syms = subprocess.run(["nm", "-DC", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    for want in ("calDisparity_SGBM(cv::Mat const&, cv::Mat const&, cv::Mat&)",
                 "triangulate10D(cv::Mat const&, cv::Mat const&, cv::Mat&, double, double, double, double, ROI3D)",
                 "correct3DPoints(cv::Mat&, ROI3D&, double const&, double const&)",
                 "setImageROI(cv::Mat&, cv::Mat&)",
                 "rgbd_tutor::Mapper::generatePointCloud(std::shared_ptr<rgbd_tutor::RGBDFrame> const&)",
                 "rgbd_tutor::Mapper::semantic_motion_fuse(std::shared_ptr<rgbd_tutor::RGBDFrame> const&)",
                 "rgbd_tutor::Mapper::viewer()", "rgbd_tutor::Mapper::SaveMap()",
                 "ssm_dropin::calUDisparity(cv::Mat const&, cv::Mat&, cv::Mat&, cv::Mat&, cv::Mat&, cv::Mat&)",
                 "ssm_dropin::calVDisparity(cv::Mat const&, cv::Mat&, cv::Mat&, cv::Mat&)"):
        assert want in syms, want
    # the reference's own FrameReader::next (compiled from /root/reference/src/rgbdframe.cpp) resolves calDisparity_SGBM to the drop-in
    undefined = subprocess.run(["nm", "-DC", "--undefined-only", path], capture_output=True, text=True, check=True).stdout
    for want in ("ssm_sgbm", "ssm_generate_point_cloud", "ssm_semantic_motion_fuse", "ssm_keyframe_add", "ssm_map_redraw", "ssm_map_export",
                 "ssm_triangulate10d", "ssm_u_disparity", "ssm_v_disparity"):
        assert want in undefined, want


def test_host_layer_compiles_with_the_cv_mat_overload(tmp_path):
    """host/host.cpp under SSM_WITH_OPENCV: void calDisparity_SGBM(const cv::Mat&, const cv::Mat&, cv::Mat&) (include/stereo.h:15)."""
    out = tmp_path / "host_cv.o"
    cmd = ["g++", "-std=c++17", "-c", "-DSSM_WITH_OPENCV", "-DCV_Assert(x)=do { if (!(x)) throw std::runtime_error(#x); } while (0)",
           "-include", "stdexcept", "-I" + os.path.join(ROOT, "oracle", "cvstub"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "semantic_slam_mapping_b200", "host", "host.cpp"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", "-C", "--defined-only", str(out)], capture_output=True, text=True, check=True).stdout
    assert "calDisparity_SGBM(cv::Mat const&, cv::Mat const&, cv::Mat&)" in syms
