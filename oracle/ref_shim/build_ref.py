#!/usr/bin/env python3
"""Build the UNMODIFIED Puresoft3D CPU renderer (read in place from /root/reference) into
oracle/_ref/libps3d_ref.so, exporting the same C-ABI as include/ps3d.h (oracle/ref_shim/ref_capi.cpp).

TEST INFRASTRUCTURE: the result is the parity pin for oracle/ps3d_oracle.c and the `--impl reference`
CPU arm of bench.py. Nothing on the product path links or loads it.

What the shim changes (no algorithmic change; BASELINE.json north_star / SURVEY.md §8c):
  * the three MSVC `__asm{}` blocks on the pipeline side are replaced by the equivalent SSE intrinsics
    (interp.cpp:55-66 haddps pair; fbo.cpp:353-370 and :377-393 16-byte fill loops);
  * src/mcemath is compiled from its own text, its asm blocks rewritten instruction by instruction (asm_translate.py);
    oracle/ref_shim/mcemath_sse.cpp (round 1: a hand restatement) is kept only as a cross-check library;
  * `typedef __declspec(align(16)) struct {...} VertexProcessorOutput` (proc.h:20-24): alignment moved
    onto the struct so gcc accepts arrays of it;
  * Win32 threads/atomics/aligned malloc come from oracle/ref_shim/include/windows.h + win32_shim.cpp;
  * GDI+/DirectDraw presenters, the picture loader and cpu.cpp are not compiled.
  * demo shaders (src/test/testproc.cpp, src/test2/testproc.cpp): their 6 asm blocks (16-byte copies and
    the FP_Cloud float->byte pack) get the same treatment.

Patched copies exist only in a scratch directory under oracle/_ref/ during the build and are deleted
afterwards; only the .so stays (git-ignored).  Flags: -O2 -msse4.1 -mmmx -mfpmath=sse -ffp-contract=off.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PS3D_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(REPO, "oracle", "_ref")
OUT_SO = os.path.join(OUT_DIR, "libps3d_ref.so")
DEMOS = os.environ.get("PS3D_REF_DEMOS", "1") == "1"   # the demos' own shader triples (src/test, src/test2)

PIPE_SOURCES = [
    "drawvao.cpp", "vertthrd.cpp", "rasterizer.cpp", "interp.cpp", "fragthrd.cpp", "fbo.cpp",
    "samplr2d.cpp", "samplrcube.cpp", "samplrproj.cpp", "proc.cpp", "pipeline.cpp", "tex.cpp", "prog.cpp",
    "vao.cpp", "vbo.cpp", "udm.cpp", "post.cpp", "dbg.cpp",
    "tex1light1.cpp", "colr1light1.cpp", "tex1bump1light1.cpp", "skybox.cpp", "shadow.cpp",
]

CXXFLAGS = [
    "-std=gnu++14", "-O2", "-fPIC", "-pthread", "-msse4.1", "-mmmx", "-mfpmath=sse", "-ffp-contract=off",
    "-fno-fast-math", "-fno-strict-aliasing", "-w",
    "-D__declspec(x)=__attribute__((x))", "-Dalign(x)=aligned(x)", "-D__stdcall=", "-D_stdcall=", "-D_cdecl=",
    "-D__cdecl=", "-D__int64=long long", "-DNDEBUG", "-D_declspec(x)=__attribute__((x))", "-D_copysign=copysign",
]

# ---- the asm replacements -------------------------------------------------------------------------------------

INTERP_ASM = """{
		/* shim: interp.cpp:55-66 - haddps x2 on each vector: (c0+c1)+(c2+c3); lane0 = left sum, lane2 = right sum */
		__m128 l_ = _mm_load_ps(correctedContributesForLeft);
		l_ = _mm_hadd_ps(l_, l_);
		l_ = _mm_hadd_ps(l_, l_);
		__m128 r_ = _mm_load_ps(correctedContributesForLeft + 4);
		r_ = _mm_hadd_ps(r_, r_);
		r_ = _mm_hadd_ps(r_, r_);
		r_ = _mm_move_ss(r_, l_);
		_mm_store_ps(temp, r_);
	}"""

FBO_CLEAR16_ASM = """{
		/* shim: fbo.cpp:353-370 - movaps fill of m_bytes/16 quads */
		__m128 v_ = _mm_load_ps((const float*)dataAligned16Bytes);
		float* p_ = (float*)m_buffer;
		for(size_t n_ = m_bytes >> 4; n_ > 0; n_--, p_ += 4)
			_mm_store_ps(p_, v_);
	}"""

FBO_CLEARZERO_ASM = """{
		/* shim: fbo.cpp:377-393 */
		__m128 v_ = _mm_setzero_ps();
		float* p_ = (float*)m_buffer;
		for(size_t n_ = m_bytes >> 4; n_ > 0; n_--, p_ += 4)
			_mm_store_ps(p_, v_);
	}"""

ASM_BLOCK = re.compile(r"__asm\s*\{.*?\n\t*\}", re.S)


def replace_asm_blocks(text, replacements, what):
    blocks = ASM_BLOCK.findall(text)
    if len(blocks) != len(replacements):
        raise SystemExit("%s: expected %d asm blocks, found %d" % (what, len(replacements), len(blocks)))
    it = iter(replacements)
    return ASM_BLOCK.sub(lambda m: next(it), text)


def patch_source(name, text):
    if name == "interp.cpp":
        text = replace_asm_blocks(text, [INTERP_ASM], name)
        text = text.replace('#include "interp.h"', '#include "interp.h"\n#include <pmmintrin.h>')
    elif name == "fbo.cpp":
        text = replace_asm_blocks(text, [FBO_CLEAR16_ASM, FBO_CLEARZERO_ASM], name)
        text = text.replace('#include "fbo.h"', '#include "fbo.h"\n#include <xmmintrin.h>')
    elif name == "proc.h":
        old = "typedef __declspec(align(16)) struct\n{\n\tfloat position[4];\n\tvoid* user;\n} VertexProcessorOutput;"
        new = "typedef struct __attribute__((aligned(16)))\n{\n\tfloat position[4];\n\tvoid* user;\n} VertexProcessorOutput;"
        text = text.replace("\r\n", "\n")
        if old not in text:
            raise SystemExit("proc.h: VertexProcessorOutput typedef not found")
        text = text.replace(old, new)
    return text


# demo shaders: src/test/testproc.cpp has 4 asm blocks, src/test2/testproc.cpp 2.
def patch_demo_shader(which, text):
    text = text.replace("\r\n", "\n")
    blocks = ASM_BLOCK.findall(text)
    out = []
    for b in blocks:
        out.append(translate_simple_asm(b, which))
    it = iter(out)
    return "#include <smmintrin.h>\n" + ASM_BLOCK.sub(lambda m: next(it), text)


def translate_simple_asm(block, what):
    """Translate the demo shaders' asm blocks, which consist only of
         mov reg, <ptr var> | lea reg, <var>        -> pointer binding
         movaps xmmN, [reg + off]                   -> _mm_load_ps
         movaps [reg + off], xmmN                   -> _mm_store_ps
         cvtps2dq / packusdw / packuswb / movd / movss / mulps / shufps ... (FP_Cloud pack)
       Anything unknown aborts the build (never guess)."""
    lines = []
    inner = block[block.index("{") + 1:block.rindex("}")]
    for raw in inner.split("\n"):
        code = raw.split(";")[0].strip()
        if code:
            lines.append(code)
    regs = {}
    body = ["{ /* shim: asm block translated 1:1 (%s) */" % what, "\t\t__m128 x_[8]; __m128i xi_[8];"]

    def addr(expr):
        expr = expr.strip()[1:-1].strip()
        m = re.match(r"(\w+)\s*(?:\+\s*(\w+))?$", expr)
        if not m:
            raise SystemExit("%s: cannot translate address %r" % (what, expr))
        off = m.group(2)
        offv = int(off, 0) if off else 0
        base = regs.get(m.group(1))
        if base is None:
            if re.match(r"e[a-ds][xi]$", m.group(1)):
                raise SystemExit("%s: register %s used before it was loaded" % (what, m.group(1)))
            base = m.group(1)  # [variable]: a C array in scope
        return "((const char*)(%s) + %d)" % (base, offv)

    for code in lines:
        m = re.match(r"(\w+)\s+(.*)$", code)
        op, args = m.group(1).lower(), [a.strip() for a in m.group(2).split(",")]
        if op in ("mov", "lea") and re.match(r"e[a-d]x|esi|edi", args[0]):
            regs[args[0]] = ("(&(%s))" % args[1]) if op == "lea" else args[1]
            if op == "lea":
                regs[args[0]] = "(%s)" % args[1]  # arrays decay; structs are passed by address below
        elif op == "movaps" and args[0].startswith("xmm") and args[1].startswith("["):
            body.append("\t\tx_[%s] = _mm_load_ps((const float*)%s);" % (args[0][3:], addr(args[1])))
        elif op == "movaps" and args[0].startswith("[") and args[1].startswith("xmm"):
            body.append("\t\t_mm_store_ps((float*)%s, x_[%s]);" % (addr(args[0]), args[1][3:]))
        elif op == "movaps" and args[0].startswith("xmm") and args[1].startswith("xmm"):
            body.append("\t\tx_[%s] = x_[%s];" % (args[0][3:], args[1][3:]))
        elif op == "cvtps2dq":
            body.append("\t\txi_[%s] = _mm_cvtps_epi32(x_[%s]);" % (args[0][3:], args[1][3:]))
        elif op == "packusdw":
            body.append("\t\txi_[%s] = _mm_packus_epi32(xi_[%s], xi_[%s]);" % (args[0][3:], args[0][3:], args[1][3:]))
        elif op == "packuswb":
            body.append("\t\txi_[%s] = _mm_packus_epi16(xi_[%s], xi_[%s]);" % (args[0][3:], args[0][3:], args[1][3:]))
        elif op == "movd" and args[0].startswith("["):
            body.append("\t\t*(int*)%s = _mm_cvtsi128_si32(xi_[%s]);" % (addr(args[0]), args[1][3:]))
        elif op == "movss" and args[0].startswith("[") and args[1].startswith("xmm"):
            # after integer packs the register holds integer data: store the low 32 bits
            body.append("\t\t*(int*)%s = _mm_cvtsi128_si32(xi_[%s]);" % (addr(args[0]), args[1][3:]))
        else:
            raise SystemExit("%s: unsupported asm statement %r" % (what, code))
    body.append("\t}")
    return "\n".join(body)


# src/test2/testpost.cpp (PP_DepthofField): two MMX asm blocks — load the constant, then movq / paddb / movntq per two pixels
# (the dead `mov eax,1 / movd mm1,eax` pair has no effect on the result).
TESTPOST_ASM = [
    "__m64 mm2_ = *(const __m64*)f; /* shim: lea eax,f / movq mm2,[eax] */",
    "{ /* shim: movq mm0,[row] / paddb mm0,mm2 / movntq [row],mm0 */\n\t\t\t\t__m64 mm0_ = *(const __m64*)row;\n\t\t\t\tmm0_ = _mm_add_pi8(mm0_, mm2_);\n\t\t\t\t_mm_stream_pi((__m64*)row, mm0_);\n\t\t\t}",
]


def patch_testpost(text):
    text = text.replace("\r\n", "\n")
    blocks = ASM_BLOCK.findall(text)
    if len(blocks) != 2:
        raise SystemExit("testpost.cpp: expected 2 asm blocks, found %d" % len(blocks))
    it = iter(TESTPOST_ASM)
    return "#include <xmmintrin.h>\n#include <mmintrin.h>\n" + ASM_BLOCK.sub(lambda m: next(it), text)


# src/mcemath: its C API (vector.cpp, matrix.cpp, quatern.cpp, matrxgl.cpp) is compiled from the reference's OWN text, every MSVC asm
# block rewritten instruction by instruction by asm_translate.py (never by hand). wraprs.cpp (C++ wrapper classes) is not on the path.
MCEMATH_SOURCES = ["vector.cpp", "matrix.cpp", "quatern.cpp", "matrxgl.cpp"]
MCEMATH_BLOCKS = {"vector.cpp": 29, "matrix.cpp": 15, "quatern.cpp": 7, "matrxgl.cpp": 1}   # live blocks (SURVEY.md §8c counts 35 / 16 / 7 / 2 with the commented-out ones)


def patch_mcemath(name, text):
    import asm_translate
    # lines that are comments from their first column on (the reference keeps a few dead asm blocks that way) go first: a
    # commented `__asm{` must not open a block
    text = "\n".join("" if ln.lstrip().startswith("//") else ln for ln in text.split("\n"))
    text, n = asm_translate.translate_source(text, "src/mcemath/" + name)
    if n != MCEMATH_BLOCKS[name]:
        raise SystemExit("src/mcemath/%s: expected %d asm blocks, translated %d" % (name, MCEMATH_BLOCKS[name], n))
    if "__asm" in text:
        raise SystemExit("src/mcemath/%s: an asm block was left untranslated" % name)
    return "#include <pmmintrin.h>\n#include <stdint.h>\n" + text


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-6000:] + "\n")
        raise SystemExit("build_ref: compile failed")
    return r.stdout


def main():
    if not os.path.isdir(os.path.join(REF, "src", "puresoft3d")):
        print("build_ref: %s not present — keeping any prebuilt %s" % (REF, OUT_SO))
        return 0
    scratch = os.path.join(OUT_DIR, "build")
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(os.path.join(scratch, "p"))
    os.makedirs(os.path.join(scratch, "t1"))
    os.makedirs(os.path.join(scratch, "t2"))
    os.makedirs(os.path.join(scratch, "m"))
    sys.path.insert(0, HERE)
    try:
        for fn in MCEMATH_SOURCES:
            with open(os.path.join(REF, "src", "mcemath", fn), "r", encoding="latin-1") as f:
                text = f.read()
            with open(os.path.join(scratch, "m", fn), "w", encoding="latin-1") as f:
                f.write(patch_mcemath(fn, text))
        src = os.path.join(REF, "src", "puresoft3d")
        for fn in sorted(os.listdir(src)):
            if fn.endswith(".h") or fn in PIPE_SOURCES:
                with open(os.path.join(src, fn), "r", encoding="latin-1") as f:
                    text = f.read()
                with open(os.path.join(scratch, "p", fn), "w", encoding="latin-1") as f:
                    f.write(patch_source(fn, text))
        for sub, dst in ((("test", "t1"), ("test2", "t2")) if DEMOS else ()):
            for fn in ("testproc.cpp", "testproc.h"):
                with open(os.path.join(REF, "src", sub, fn), "r", encoding="latin-1") as f:
                    text = f.read()
                if fn.endswith(".cpp"):
                    text = patch_demo_shader("src/%s/%s" % (sub, fn), text)
                with open(os.path.join(scratch, dst, fn), "w", encoding="latin-1") as f:
                    f.write(text)
        if DEMOS:
            for fn in ("testpost.cpp", "testpost.h"):
                with open(os.path.join(REF, "src", "test2", fn), "r", encoding="latin-1") as f:
                    text = f.read()
                with open(os.path.join(scratch, "t2", fn), "w", encoding="latin-1") as f:
                    f.write(patch_testpost(text) if fn.endswith(".cpp") else text)
        inc = ["-I", os.path.join(HERE, "include"), "-I", os.path.join(scratch, "p"),
               "-I", os.path.join(REF, "src", "mcemath"), "-I", os.path.join(REPO, "include"),
               "-include", os.path.join(HERE, "prelude.h")]
        objs = []
        jobs = [(os.path.join(scratch, "p", s), "p_" + s) for s in PIPE_SOURCES]
        jobs += [] if not DEMOS else [(os.path.join(scratch, "t1", "testproc.cpp"), "t1_testproc.cpp"),
                 (os.path.join(scratch, "t2", "testproc.cpp"), "t2_testproc.cpp"), (os.path.join(scratch, "t2", "testpost.cpp"), "t2_testpost.cpp"),
                 (os.path.join(HERE, "demo_procs1.cpp"), "t1_demo_procs1.cpp"), (os.path.join(HERE, "demo_procs2.cpp"), "t2_demo_procs2.cpp")]
        jobs += [(os.path.join(scratch, "m", s), "m_" + s) for s in MCEMATH_SOURCES]
        jobs += [(os.path.join(HERE, s), "s_" + s) for s in ("win32_shim.cpp", "ref_capi.cpp")]
        for path, tag in jobs:
            obj = os.path.join(scratch, tag + ".o")
            extra = []
            if tag.startswith("t1_"):
                extra = ["-I", os.path.join(scratch, "t1"), "-DPS3D_DEMO=1"]
            if tag.startswith("t2_"):
                extra = ["-I", os.path.join(scratch, "t2"), "-DPS3D_DEMO=2"]
            if tag == "s_ref_capi.cpp" and DEMOS:
                extra = ["-DPS3D_WITH_DEMO_SHADERS=1"]
            run(["g++"] + CXXFLAGS + inc + extra + ["-c", path, "-o", obj])
            objs.append(obj)
        os.makedirs(OUT_DIR, exist_ok=True)
        run(["g++", "-shared", "-pthread", "-o", OUT_SO] + objs)
        print("build_ref: wrote", OUT_SO)
        # the hand-written restatement of the routines the pipeline calls (round 1's shim) stays as a CROSS-CHECK of the mechanical
        # translation: its own library, compared routine by routine by tests/test_mcemath_translation.py
        hand = os.path.join(OUT_DIR, "libmcemath_hand.so")
        run(["g++"] + CXXFLAGS + ["-shared", "-o", hand, os.path.join(HERE, "mcemath_sse.cpp")])
        print("build_ref: wrote", hand)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
