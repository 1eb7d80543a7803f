"""Stage the reference's OWN wrappers where the GPU box can run them against our library.

/root/reference exists only in the build container.  The drop-in claim -- the unchanged Python ctypes wrapper
(python/dsc/*.py, _bindings.py:31-35 loads ./libdsc.so) and the unchanged C++ header wrapper (dsc/api/dsc_api.h)
work on top of dsc_b200/libdsc.so -- has to be shown on the B200, so this script copies them into
baseline/_ref/ (git-ignored, never committed, but part of the gpurun snapshot):

    baseline/_ref/python/dsc/*.py            the reference wrapper, byte for byte
    baseline/_ref/python/tests/test_ops.py   the reference's own test-suite (16 tests, FFT included)
    baseline/_ref/cpp/filter_readme          README.md:118-134 through dsc_api.h, compiled HERE against the
                                             reference's dsc_api.h + our include/dsc.h, linked to dsc_b200/libdsc.so

tests/test_dropin_gpu.py (-m gpu) runs both on the box.  Run by __graft_entry__.build() when /root/reference exists.
"""
import os
import shutil
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DSC_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

README_FILTER = textwrap.dedent("""
    #include "dsc_api.h"
    #include <cmath>
    #include <cstdio>
    #include <vector>
    int main() {
        dsc::init(1 << 28);
        const int n = 8192, taps = 128, fft_size = 16384;
        std::vector<f32> sv(n), bv(taps, 1.f / taps);
        for (int i = 0; i < n; ++i) sv[i] = std::sin(0.01f * i);
        dsc::tensor<f32> s(sv.data(), n), b(bv.data(), taps);
        dsc::tensor<f32> S = dsc::rfft(s, fft_size);
        dsc::tensor<f32> B = dsc::rfft(b, fft_size);
        dsc::tensor<f32> conv = S * B;
        dsc::tensor<f32> y = dsc::irfft(conv);
        dsc::tensor<f32> out = y.get(DSC_SLICE_TO(n + taps - 1));
        if (out.size() != n + taps - 1) { printf("bad size %d\\n", out.size()); return 1; }
        double err = 0;
        for (int i = taps; i < n; ++i) {
            double want = 0;
            for (int k = 0; k < taps; ++k) want += sv[i - k] / taps;
            err = std::fmax(err, std::fabs(out.data()[i] - want));
        }
        printf("samples=%d max_err=%.3g\\n", out.size(), err);
        return err < 1e-4 ? 0 : 2;
    }
""")


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "python", "dsc")):
        print(f"stage_reference_wrappers: {REF} not present, nothing staged")
        return 0
    py_dst = os.path.join(DST, "python")
    shutil.rmtree(py_dst, ignore_errors=True)
    os.makedirs(os.path.join(py_dst, "dsc"))
    os.makedirs(os.path.join(py_dst, "tests"))
    for f in os.listdir(os.path.join(REF, "python", "dsc")):
        if f.endswith(".py") or f == "py.typed":
            shutil.copyfile(os.path.join(REF, "python", "dsc", f), os.path.join(py_dst, "dsc", f))
    shutil.copyfile(os.path.join(REF, "python", "tests", "test_ops.py"), os.path.join(py_dst, "tests", "test_ops.py"))

    cpp_dst = os.path.join(DST, "cpp")
    os.makedirs(cpp_dst, exist_ok=True)
    lib = os.path.join(ROOT, "dsc_b200", "libdsc.so")
    if os.path.exists(lib):
        src = os.path.join(cpp_dst, "filter_readme.cpp")
        with open(src, "w") as fh:
            fh.write(README_FILTER)
        cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
        cmd = ["g++", "-std=c++20", "-O1", f"-I{ROOT}/include", f"-I{REF}/dsc/api", src, "-o", os.path.join(cpp_dst, "filter_readme"),
               f"-L{os.path.dirname(lib)}", "-ldsc", "-Wl,-rpath,$ORIGIN/../../../dsc_b200", f"-Wl,-rpath-link,{cuda_lib}", "-pthread"]
        subprocess.run(cmd, check=True)
        os.remove(src)          # the binary is what travels; the source is in this script
    print(f"stage_reference_wrappers: staged under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
