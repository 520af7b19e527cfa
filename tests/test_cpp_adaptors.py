"""The header-only openMVG::matching::ArrayMatcher adaptor (regard3d_b200/csrc/ArrayMatcher_b200.h) compiles against a
stand-in of the two OpenMVG types it is instantiated with, links with libr3dgpu.so, and -- on a machine without a
B200 -- fails loudly (Build returns false, the error text says there is no CPU fallback)."""
import os
import subprocess
import textwrap

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

SRC = textwrap.dedent(r'''
    #include <cstdio>
    #include <vector>
    #include "regard3d_b200/csrc/ArrayMatcher_b200.h"
    // stand-ins with the shape of openMVG/matching/matching_interface.hpp, metric.hpp, indMatch.hpp
    namespace openMVG { namespace matching {
    struct IndMatch { IndMatch(uint32_t i = 0, uint32_t j = 0) : i_(i), j_(j) {} uint32_t i_, j_; };
    using IndMatches = std::vector<IndMatch>;
    template <typename T> struct L2 { typedef T ElementType; typedef float ResultType; };
    template <typename Scalar, typename Metric> class ArrayMatcher {
     public:
      using ScalarT = Scalar; using DistanceType = typename Metric::ResultType;
      virtual ~ArrayMatcher() = default;
      virtual bool Build(const Scalar* dataset, int nbRows, int dimension) = 0;
      virtual bool SearchNeighbour(const Scalar* query, int* indice, DistanceType* distance) = 0;
      virtual bool SearchNeighbours(const Scalar* query, int nbQuery, IndMatches* indices,
                                    std::vector<DistanceType>* distances, size_t NN) = 0;
    };
    }}
    using namespace openMVG::matching;
    using GpuMatcher = r3d_shim::ArrayMatcher_b200<float, L2<float>, ArrayMatcher<float, L2<float>>, IndMatch>;
    int main() {
      GpuMatcher m;
      ArrayMatcher<float, L2<float>>* base = &m;   // usable through the OpenMVG interface
      std::vector<float> db(8 * 4, 0.5f), q(2 * 4, 0.25f);
      const bool built = base->Build(db.data(), 8, 4);
      IndMatches idx; std::vector<float> dist;
      const bool ok = built && base->SearchNeighbours(q.data(), 2, &idx, &dist, 2);
      std::printf("built=%d searched=%d n=%zu err=%s\n", (int)built, (int)ok, idx.size(), m.lastError().c_str());
      return 0;
    }
''')


def _build_and_run(tmp_path):
    src = tmp_path / "adaptor.cpp"
    src.write_text(SRC)
    exe = tmp_path / "adaptor"
    libdir = os.path.join(ROOT, "regard3d_b200")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-I", ROOT, str(src), "-o", str(exe), "-L", libdir, "-lr3dgpu",
           "-Wl,-rpath," + libdir]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_array_matcher_adaptor_compiles_links_and_fails_loudly_without_gpu(r3dlib, tmp_path):
    out = _build_and_run(tmp_path)
    import torch
    if torch.cuda.is_available():
        assert "built=1 searched=1 n=4" in out
    else:
        assert "built=0 searched=0" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_array_matcher_adaptor_runs_on_the_gpu(r3dlib, tmp_path):
    """The adaptor's GPU branch (Build -> r3d_upload_regions, SearchNeighbours -> r3d_search_neighbours) on hardware."""
    assert "built=1 searched=1 n=4" in _build_and_run(tmp_path)
