"""The C++ adapters compile against the reference's OWN headers and types: tests/native/adapter_signatures.cpp
includes /root/reference/include/slam/{common,frontend}/*.hpp -- through oracle/shim's Eigen / OpenCV / spdlog stand-ins -- and
instantiates every adapter method with slam::EigenGrayMatrix, slam::Keypoint, slam::DescriptorMatrix, slam::Match,
slam::KeyDescriptorPair, slam::Camera, cv::Mat, cv::KeyPoint, cv::DMatch, cv::Point3d, Eigen::MatrixXd.  Compile only."""
import os
import subprocess

import pytest

from conftest import ROOT

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "slam")), reason="/root/reference is not here (GPU box)")
def test_adapters_accept_the_reference_types(tmp_path):
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-w", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(REF, "include"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "native", "adapter_signatures.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


def test_adapters_compile_standalone(tmp_path):
    """... and with this repository's Eigen-free look-alike types (the host programs the GPU tests run)."""
    for src in ("tools/cli/slam_bench.cpp", "test/frontend/test_frontend_cuda.cpp"):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, src)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
