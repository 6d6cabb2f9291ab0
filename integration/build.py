"""Builds integration/_build/gs_ref_binding*.so: the reference's pybind module (rasterize_gaussians /
rasterize_gaussians_backward / mark_visible) on top of libgsplat_b200.so.  No GPU needed to build.
usage: python integration/build.py   (also called by __graft_entry__.build())"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "gaussian-pcloud-render_b200")
NAME = "gs_ref_binding"


def build(verbose: bool = False):
    from torch.utils.cpp_extension import load
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    if not os.path.exists(os.path.join(PKG, "libgsplat_b200.so")):
        raise RuntimeError("build the library first: make -C gaussian-pcloud-render_b200/csrc")
    return load(name=NAME, sources=[os.path.join(HERE, "rasterize_points_b200.cpp")],
                extra_include_paths=[os.path.join(ROOT, "include")], extra_cflags=["-O2"],
                extra_ldflags=[f"-L{PKG}", "-lgsplat_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../gaussian-pcloud-render_b200"],
                build_directory=out, with_cuda=True, verbose=verbose)


def load_built():
    """Imports the already built module from integration/_build (no compilation; raises if it is missing)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    path = os.path.join(HERE, "_build", NAME + ".so")
    if not os.path.exists(path):
        raise ImportError(f"{path} missing: run python integration/build.py")
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = build(verbose="-v" in sys.argv)
    print("built", m.__file__)
