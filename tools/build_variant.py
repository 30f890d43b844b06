"""Developer A/B builds of the row kernel: compile conv_row.cu / conv_row_cp.cu with extra -D flags and link them with the
product objects into larvanet_b200/csrc/build/variants/lib_<name>.so; select one with LARVANET_B200_LIB=<path>.

    python tools/build_variant.py NAME -DLV_ROW_FOO=1 [...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from larvanet_b200 import build as B  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    B.build()
    vdir = os.path.join(B.BUILD, 'variants')
    os.makedirs(vdir, exist_ok=True)
    objs = []
    procs = []
    for s in B.SOURCES:
        o = os.path.join(B.BUILD, s.replace('.cu', '.o'))
        if s.startswith('conv_row') or s.startswith('conv_chain'):
            o = os.path.join(vdir, f'{name}_{s.replace(".cu", ".o")}')
            procs.append(subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + defs + ['-c', os.path.join(B.CSRC, s), '-o', o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        objs.append(o)
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise SystemExit(1)
    lib = os.path.join(vdir, f'lib_{name}.so')
    subprocess.check_call([B._nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'])
    print(lib)


if __name__ == '__main__':
    main()
