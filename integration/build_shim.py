"""Build integration/libmpxb_xla.so: the XLA-FFI adapter (xla_ffi_shim.cc) over libmpxb.so.

Against the real header when JAX is installed (jax.ffi.include_dir()), else against the minimal stand-in under
integration/stub/ (TEST INFRASTRUCTURE, see its header) so that the shim source is compiled and exercised here."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libmpxb_xla.so")


def include_dir():
    try:
        import jax  # noqa: F401
        from jax import ffi
        return ffi.include_dir(), True
    except Exception:
        return os.path.join(HERE, "stub"), False


def build(force=False):
    src = os.path.join(HERE, "xla_ffi_shim.cc")
    inc, real = include_dir()
    deps = [src, os.path.join(ROOT, "include", "mpxb.h"), os.path.join(ROOT, "mpx_b200", "libmpxb.so")]
    if not real:
        deps.append(os.path.join(inc, "xla", "ffi", "api", "ffi.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps if os.path.exists(d)):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I" + inc, "-I" + os.path.join(ROOT, "include"), src,
           "-L" + os.path.join(ROOT, "mpx_b200"), "-lmpxb", "-Wl,-rpath," + os.path.join(ROOT, "mpx_b200"), "-o", OUT]
    if real:
        cmd.insert(1, "-I/usr/local/cuda/include")
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
