"""In-tree build of the CUDA library (sm_100a only).  `python -m elisa_b200.build`."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'elisa_b200.cu')
DEPS = [os.path.join(HERE, 'csrc', f) for f in ('elisa_b200.cu', 'components.cuh', 'dual.cuh', 'fastmath.cuh', 'fold_i8.cuh', 'gemm_f64.cuh', 'slice_i8.cuh', 'stats.cuh', 'unfold_skel.cuh')]
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
DEPS.append(os.path.join(INCLUDE, 'elisa_b200.h'))
# headers handed to NVRTC at run time travel inside the library as string literals
EMBED = {
    'embedded_elisa_b200_h.inc': os.path.join(INCLUDE, 'elisa_b200.h'),
    'embedded_dual_cuh.inc': os.path.join(HERE, 'csrc', 'dual.cuh'),
    'embedded_components_cuh.inc': os.path.join(HERE, 'csrc', 'components.cuh'),
    'embedded_unfold_skel_cuh.inc': os.path.join(HERE, 'csrc', 'unfold_skel.cuh'),
    'embedded_slice_i8_cuh.inc': os.path.join(HERE, 'csrc', 'slice_i8.cuh'),
    'embedded_fastmath_cuh.inc': os.path.join(HERE, 'csrc', 'fastmath.cuh'),
    'embedded_stats_cuh.inc': os.path.join(HERE, 'csrc', 'stats.cuh'),
}
LIB = os.path.join(HERE, 'lib', 'libelisa_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared', '-I', INCLUDE,
]


def embed_sources() -> None:
    """csrc/*.cuh -> csrc/_gen/embedded_*.inc (C++ raw string literals, split below the
    compilers' per-literal limit)."""
    out_dir = os.path.join(HERE, 'csrc', '_gen')
    os.makedirs(out_dir, exist_ok=True)
    for name, path in EMBED.items():
        text = open(path).read()
        chunks = [text[i:i + 8000] for i in range(0, len(text), 8000)]
        body = '\n'.join(f'R"ELISASRC({c})ELISASRC"' for c in chunks) + '\n'
        dst = os.path.join(out_dir, name)
        if not os.path.exists(dst) or open(dst).read() != body:
            with open(dst, 'w') as f:
                f.write(body)


def find_nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: elisa_b200 needs the CUDA toolkit to build its sm_100a library')
    return nvcc


def is_fresh() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/elisa_b200.cu -> lib/libelisa_b200.so (skipped when up to date)."""
    if not force and is_fresh():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    embed_sources()
    cmd = [find_nvcc(), *NVCC_FLAGS, '-I', os.path.join(HERE, 'csrc', '_gen'), '-o', LIB, SRC, '-ldl']
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libelisa_b200.so')
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


XLA_SRC = os.path.join(HERE, 'csrc', 'elisa_b200_xla.cc')
XLA_LIB = os.path.join(HERE, 'lib', 'libelisa_b200_xla.so')


def xla_ffi_include_dir() -> str | None:
    """Where xla/ffi/api/ffi.h lives (jax.ffi.include_dir()), or None when JAX is not installed --
    the case in the image this library is developed in."""
    try:
        import jax  # noqa: F401
        from jax import ffi as jffi

        d = jffi.include_dir()
    except Exception:
        return None
    return d if os.path.exists(os.path.join(d, 'xla', 'ffi', 'api', 'ffi.h')) else None


def build_xla_shim(force: bool = False) -> str | None:
    """Compile csrc/elisa_b200_xla.cc -> lib/libelisa_b200_xla.so (the jax.ffi handlers of
    INTEGRATION.md section 4) when the XLA FFI headers are present; returns None (and builds
    nothing) when they are not."""
    inc = xla_ffi_include_dir()
    if inc is None:
        return None
    if not force and os.path.exists(XLA_LIB) and os.path.getmtime(XLA_LIB) >= max(os.path.getmtime(XLA_SRC), os.path.getmtime(LIB)):
        return XLA_LIB
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-I', inc, '-I', INCLUDE, '-o', XLA_LIB, XLA_SRC,
           '-L', os.path.dirname(LIB), '-lelisa_b200', '-Wl,-rpath,$ORIGIN']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('building the XLA FFI shim failed')
    return XLA_LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    shim = build_xla_shim(force='--force' in sys.argv)
    print(shim if shim else 'XLA FFI headers not found (jax absent): csrc/elisa_b200_xla.cc not built')
