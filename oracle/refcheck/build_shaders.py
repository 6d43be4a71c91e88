"""oracle/refcheck/build_shaders.py -- TEST INFRASTRUCTURE.

Compiles the REFERENCE's own fragment shaders, from where they lie under /root/reference, as C++ on
top of the reference's vendored glm (see glsl_prelude.h), into oracle/_ref/libvxshader.so.  This is
"the reference itself run here" for the ray-march functions and the ray generation of the four light
passes; tests/test_oracle_shaders.py pins the oracle (oracle/vxo.cpp) against it bit for bit.

No reference source is copied into the repository: the shader text is read at build time, passed
through the substitutions listed in SUBSTITUTIONS (each one a syntactic adaptation from GLSL to
C++ that leaves every arithmetic expression untouched) and piped to g++ on stdin.  The only build
product is the shared library under oracle/_ref/ (git-ignored).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref", "libvxshader.so")
REF = os.environ.get("VXL_REFERENCE", "/root/reference")
SH = os.path.join(REF, "Sources", "Shaders")

# (regex, replacement, why)
SUBSTITUTIONS = [
    (r"^#version .*$", r"// \g<0>", "GLSL-only directive"),
    (r"^#extension .*$", r"// \g<0>", "GLSL-only directive"),
    (r'^#include "(lib/)?Common.frag"\s*$', r"// \g<0>", "Common.frag is spliced once at global scope"),
    (r"\bout vec3\b", r"vec3&", "GLSL out-parameter -> C++ reference"),
    (r"\bout uint\b", r"uint&", "GLSL out-parameter -> C++ reference"),
    (r"layout\(push_constant\) uniform uPushConstant\s*\{", r'extern "C++" {', "push-constant block members become plain globals"),
    # the two march functions are renamed so that the prelude's wrappers (same name, same signature) can log each call
    (r"^float raycastShadowVolume(Sparse|SuperSparse)\(vec3 origin, vec3 dir, float dist\) \{",
     r"float raycastShadowVolume\1_impl(vec3 origin, vec3 dir, float dist) {", "call logging hook"),
    # GLSL constructors consume only as many components as they need (vec4(vec3, vec3) takes the second one's .x); colour output only
    (r"vec4\(ambient\*F\*\(1\.0-roughness\), F\)", r"vec4(ambient*F*(1.0-roughness), (F).x)", "GLSL constructor truncation"),
    # writes through a swizzle (LightTAA.frag:123,124,129): glm's function-style swizzles return values, so the statement is
    # re-spelt on the whole vec4 with the same right-hand side
    (r"^(\s*)(\w+)\.xyz\s*([+/])=\s*([^;]+);", r"\1\2 = vec4(vec3(\2) \3 (\4), \2.w);", "swizzle compound assignment"),
    (r"^(\s*)(\w+)\.xyz\s*=\s*([^;=][^;]*);", r"\1\2 = vec4(\3, \2.w);", "swizzle assignment"),
    # LightTAA.frag:30-35 names a parameter `rgb`, which collides with the swizzle macro of the prelude
    (r"\(vec3 rgb\)", r"(vec3 rgb_)", "parameter name vs swizzle macro"),
    (r"dot\(rgb, W\)", r"dot(rgb_, W)", "parameter name vs swizzle macro"),
]


def adapt(text: str) -> str:
    for rx, rep, _ in SUBSTITUTIONS:
        text = re.sub(rx, rep, text, flags=re.M)
    return text


def read(rel: str) -> str:
    with open(os.path.join(SH, rel), "r", encoding="utf-8", errors="replace") as f:
        return f.read().replace("\r\n", "\n")


def splice_includes(text: str) -> str:
    def rep(m):
        return "\n// ---- begin %s ----\n%s\n// ---- end %s ----\n" % (m.group(1), adapt(read(m.group(1))), m.group(1))
    return re.sub(r'^#include "(lib/(?:PBR|Light)\.frag)"\s*$', rep, text, flags=re.M)


WRAPPERS = """
float raycastShadowVolumeSparse_impl(vec3 origin, vec3 dir, float dist);
float raycastShadowVolumeSuperSparse_impl(vec3 origin, vec3 dir, float dist);
inline float raycastShadowVolumeSparse(vec3 origin, vec3 dir, float dist) { return vxref::logged(0, origin, dir, dist, raycastShadowVolumeSparse_impl); }
inline float raycastShadowVolumeSuperSparse(vec3 origin, vec3 dir, float dist) { return vxref::logged(1, origin, dir, dist, raycastShadowVolumeSuperSparse_impl); }
"""

PASSES = [("ambient", "LightAmbient.frag", "ViewBuffer_t ViewBuffer[1];"),
          ("point", "LightPoint.frag", "ViewBuffer_t ViewBuffer[1]; PointLightsBuffer_t PointLightsBuffer[1];"),
          ("spot", "LightSpot.frag", "ViewBuffer_t ViewBuffer[1]; SpotLightsBuffer_t SpotLightsBuffer[1];"),
          ("reflection", "LightReflection.frag", "ViewBuffer_t ViewBuffer[1];")]


def translation_unit() -> str:
    parts = ['#include "glsl_prelude.h"\n#include "shader_log.h"\n#define in\n#define out\n']
    parts.append("// ==== lib/Common.frag (global scope) ====\n" + adapt(read("lib/Common.frag")) + "\n")
    parts.append("sampler2D _BindingSampler2D[16]; usampler3D _BindingUSampler3D[4]; samplerCube _BindingSamplerCube[4];\n")
    # Light.frag on its own (no IMPORT: it declares _ShadowVoxRID = 0 itself): the ray-level entry
    parts.append("namespace light_only {\n" + WRAPPERS + adapt(read("lib/Light.frag")) + "\n}\n")
    # the vertex shader's ComputeFarVec (LightAmbient.vert:32-36), evaluated per pixel
    parts.append("namespace ambient_vert {\nint gl_VertexIndex; vec4 gl_Position;\n" + adapt(read("LightAmbient.vert")) +
                 "\nViewBuffer_t ViewBuffer[1];\n}\n")
    for ns, fn, defs in PASSES:
        parts.append("namespace %s {\n%s\n%s\n%s\n}\n" % (ns, WRAPPERS, splice_includes(adapt(read(fn))), defs))
    # the G-buffer producer: GeometryVoxel.frag -- clipToAABB, intersectVolume and the fragment main()
    geom = adapt(read("GeometryVoxel.frag"))
    # the step after the light passes: LightTAA.frag
    parts.append("namespace taa {\n%s\nViewBuffer_t ViewBuffer[1];\n}\n" % adapt(read("LightTAA.frag")))
    parts.append("namespace geom {\nfloat gl_FragDepth;\n" + geom + "\nViewBuffer_t ViewBuffer[1]; VoxCmdsBuffer_t VoxCmdsBuffer[1];\n}\n")
    with open(os.path.join(HERE, "shader_driver.inc"), "r") as f:
        parts.append(f.read())
    return "".join(parts)


def build(verbose: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REF, "Vendor", "glm")) or not os.path.isdir(SH):
        return None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [os.path.join(HERE, f) for f in ("build_shaders.py", "glsl_prelude.h", "shader_log.h", "shader_driver.inc")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-x", "c++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-fopenmp", "-fsingle-precision-constant", "-ffp-contract=off",
           "-fno-fast-math", "-I" + os.path.join(REF, "Vendor"), "-I" + HERE, "-o", OUT, "-"]
    tu = translation_unit()
    if verbose:
        sys.stderr.write("\n".join("%5d  %s" % (i + 1, l) for i, l in enumerate(tu.split("\n"))) + "\n")
    r = subprocess.run(cmd, input=tu, text=True, capture_output=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[:20000])
        raise RuntimeError("g++ failed on the adapted reference shaders")
    return OUT


if __name__ == "__main__":
    print(build(verbose="--dump" in sys.argv))
