#!/bin/bash
# ORACLE / reference build (test infrastructure only): compiles the REFERENCE'S OWN sources, where they lie under $REF, into
# oracle/_ref/libbtrcrefcpu.so together with the host shims of oracle/refcpu/.  Nothing is copied into the repo: objects and
# the one generated header go to oracle/_ref/ (git-ignored; the .so travels to the GPU box like our own built libraries).
#   * CUJ (lib/cuj/cuj/src) without its LLVM back ends (gen/llvm*, mcjit, ptx, nvrtc need LLVM 11.1 / NVRTC);
#   * src/common, src/core, src/factory, src/builtin except the files that need the OptiX SDK (utils/optix/{context,sbt,pipeline}.cpp,
#     wavefront/{trace,shadow}.cpp, renderer/wavefront.cpp, renderer/pt*, postprocess/optix_denoiser.cpp), builtin/register.cpp
#     (names those renderers) and utils/file.cpp (a bare `#elif` gcc rejects; its read_txt_file is restated in refcpu.cpp);
#   * one header is patched on the way in: cuj/dsl/class.h declares `template<> struct MemberInfo<N>` at class scope
#     (class.h:288-289), which gcc rejects (explicit specialisation in non-namespace scope); the copy under
#     oracle/_ref/gen_inc makes it a partial specialisation with a defaulted dummy parameter -- same meaning, 4 lines changed by sed.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/../_ref
OBJ=$OUT/obj_refcpu
CXX=/usr/bin/g++
test -d $REF/src/builtin || { echo "reference tree not present; keeping prebuilt _ref"; exit 1; }
mkdir -p $OBJ $OUT/gen_inc/cuj/dsl
CLS=$REF/lib/cuj/cuj/inc/cuj/dsl/class.h
python3 - "$CLS" "$OUT/gen_inc/cuj/dsl/class.h" <<'PY'
import sys
a = open(sys.argv[1]).read().split("\n")
b = list(a)
def pad(new, old):  # keep the continuation backslash in its column
    return new + " " * max(1, len(old.rstrip("\\").rstrip()) + (len(old) - len(old.rstrip("\\").rstrip()) - 1) - len(new)) + "\\"
for i, l in enumerate(a):
    nxt = a[i + 1] if i + 1 < len(a) else ""
    if l.strip().rstrip("\\").strip() == "template<int N>" and nxt.strip().rstrip("\\").strip() == "struct MemberInfo":
        b[i] = pad("        template<int N, typename CujDummy = void>", l)      # class-scope primary templates (class.h:253, :351)
    if l.strip().rstrip("\\").strip() == "template<>" and "struct MemberInfo<cuj_member_counter##NAME>" in nxt:
        b[i] = pad("    template<typename CujDummy>", l)                          # class.h:288-289
        b[i + 1] = pad("    struct MemberInfo<cuj_member_counter##NAME, CujDummy>", nxt)
d = [i + 1 for i, (x, y) in enumerate(zip(a, b)) if x != y]
assert len(d) == 4, d
open(sys.argv[2], "w").write("\n".join(b))
PY
INC="-I$OUT/gen_inc -I$HERE/stubs -I$REF/src/common -I$REF/src/core -I$REF/src/factory -I$REF/src/builtin -I$REF/lib/cuj/cuj/inc -I$REF/lib/fmt/include \
 -I$REF/lib -I$REF/lib/cpp-member-function-trait -I$REF/lib/tinyobjloader -I$REF/lib/nlohmann -I$REF/lib/stb -I$REF/lib/tinyexr -I$REF/lib/tinyexr/deps/miniz \
 -I$REF/lib/bvh/include -I/usr/local/cuda/include -DFMT_HEADER_ONLY"
FLAGS="-std=c++20 -O2 -ffp-contract=off -fopenmp -fPIC -w -include limits -include cmath -include cstring -include atomic -include $HERE/../ref_stubs/msvc_compat.h"
SRCS=$( (cd $REF/src && find common core builtin factory -name '*.cpp' | grep -v -e 'utils/optix/context.cpp' -e 'utils/optix/sbt.cpp' \
          -e 'utils/optix/pipeline.cpp' -e 'utils/file.cpp' -e 'renderer/pt' -e 'wavefront/trace.cpp' -e 'wavefront/shadow.cpp' \
          -e 'renderer/wavefront.cpp' -e 'optix_denoiser' -e 'builtin/register.cpp' | sed "s|^|$REF/src/|");
        (find $REF/lib/cuj/cuj/src -name '*.cpp' | grep -v -e 'gen/llvm' -e 'gen/mcjit' -e 'gen/ptx.cpp' -e 'gen/nvrtc');
        echo $REF/lib/tinyobjloader/tiny_obj_loader.cc $REF/lib/stb/stb/stb.cpp $REF/lib/tinyexr/tinyexr.cc $REF/lib/tinyexr/deps/miniz/miniz.c;
        ls $HERE/*.cpp )
compile_one() {
    src=$1; o=$OBJ/$(echo "$src" | sed -e "s|$REF/||" -e "s|$HERE/|refcpu_|" | tr '/' '_').o
    if [ ! -f "$o" ] || [ "$src" -nt "$o" ] || [ "$HERE/refcpu_internal.h" -nt "$o" -a "${src#$HERE}" != "$src" ]; then
        case "$src" in *.c) /usr/bin/gcc -O2 -fPIC -w -c "$src" -o "$o";; *) $CXX $FLAGS $INC -c "$src" -o "$o";; esac || { echo "FAILED: $src"; exit 1; }
    fi
}
export -f compile_one; export REF HERE OBJ CXX FLAGS INC
echo "$SRCS" | tr ' ' '\n' | grep -v '^$' | xargs -P ${JOBS:-8} -I{} bash -c 'compile_one {}'
$CXX -shared -fopenmp -Wl,-Bsymbolic -o $OUT/libbtrcrefcpu.so $OBJ/*.o -ldl
echo "built $OUT/libbtrcrefcpu.so"
