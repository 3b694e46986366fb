"""ORACLE — TEST INFRASTRUCTURE ONLY.
Does by hand what the reference's CMake configure step does for ONE file: fills the @VARIABLES@ of
modules/utils/include/litiv/utils/defines.hpp.in (read where it lies under the reference tree) with the values of a default
Linux x86-64 release configuration without GLSL / CUDA / OpenGM / Boost, and writes oracle/_ref/include/litiv/utils/defines.hpp.
Nothing of the reference is copied into the repository: the output lives under the git-ignored oracle/_ref/.
usage: configure_defines.py <reference root> <output file>"""
import os
import re
import sys

VALUES = {
    "LITIV_VERSION": "1.6.0", "LITIV_VERSION_MAJOR": "1", "LITIV_VERSION_MINOR": "6", "LITIV_VERSION_PATCH": "0", "GIT_SHA1": "unknown",
    "USE_PROFILING": "0", "USE_OPENCV_MAT_CONSTR_FIX": "0", "USE_OPENCV_x264_TEST": "0", "BUILD_TESTS_FULL_FLOAT": "0",
    "USE_SIGNEXT_SHIFT_TRICK": "0", "USE_FAST_SQRT_FOR_CDIST": "0",       # modules/utils/CMakeLists.txt:19-20 (both OFF)
    "USE_BSDS500_BENCHMARK": "0", "USE_KINECTSDK_STANDALONE": "0", "USE_CVCORE_WITH_UTILS": "1",   # CMakeLists.txt:94 (ON)
    "EXTERNAL_DATA_ROOT": "/tmp/litiv_data", "SAMPLES_DATA_ROOT": "/tmp/litiv_samples", "TEST_INPUT_DATA_ROOT": "/tmp/litiv_test_in",
    "TEST_OUTPUT_DATA_ROOT": "/tmp/litiv_test_out", "DATASETS_CACHE_SIZE": "512", "USE_RLUTIL_ANSI_DEFINE": "0", "RLUTIL_STRING_TYPE": "std::string",
    "TARGET_PLATFORM_x64": "1", "BUILD_SHARED_LIBS": "1", "USE_LINK_TIME_OPTIM": "0", "USE_FAST_MATH": "0",                # CMakeLists.txt:97 (OFF)
    "USE_OPENMP": "0", "USE_WORLD_SOURCE_GLOB": "0", "USE_SOSPD": "0", "USE_OFDIS": "0", "USE_LZ4": "0", "USE_VERSION_TAGS": "0",
    "USE_GLSL": "0", "TARGET_GL_VER_MAJOR": "4", "TARGET_GL_VER_MINOR": "4", "USE_GLEW_EXPERIMENTAL": "0", "USE_GLFW": "0", "USE_FREEGLUT": "0",
    "USE_VPTZ_STANDALONE": "0", "USE_CUDA": "0", "CUDA_VERSION_MAJOR": "0", "CUDA_VERSION_MINOR": "0", "CUDA_VERSION": "0", "CUDA_64_BIT_DEVICE_CODE": "1",
    "CUDA_PROPAGATE_HOST_FLAGS": "0", "USE_OPENCL": "0", "USE_BOOST": "0", "USE_OPENGM": "0", "USE_KINECTSDK": "0",
    # SIMD: what the reference's CMake detects on an x86-64-v3 host (the LBSP threshold takes its SSE2 branch, LBSP.hpp:203-223)
    "USE_NEON": "0", "USE_MMX": "1", "USE_SSE": "1", "USE_SSE2": "1", "USE_SSE3": "1", "USE_SSSE3": "1", "USE_SSE4_1": "1", "USE_SSE4_2": "1",
    "USE_POPCNT": "1", "USE_AVX": "1", "USE_AVX2": "1", "USE_STL_ALIGNED_ALLOC": "1", "USE_POSIX_ALIGNED_ALLOC": "1",
}


def main():
    ref, out = sys.argv[1], sys.argv[2]
    src = open(os.path.join(ref, "modules/utils/include/litiv/utils/defines.hpp.in")).read()

    def sub(m):
        k = m.group(1)
        if k.startswith(("Boost_", "USE_OPENGM_WITH_", "HAVE_OPENGM_")):
            return "0"  # sub-options of packages that are switched off above
        if k not in VALUES:
            raise SystemExit(f"configure_defines.py: no value for @{k}@")
        return VALUES[k]
    txt = re.sub(r"@([A-Za-z0-9_]+)@", sub, src)
    txt = txt.replace("#endif / unknown platform?", "#endif // unknown platform?")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write(txt)


if __name__ == "__main__":
    main()
