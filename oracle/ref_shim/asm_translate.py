"""Mechanical translation of MSVC `__asm { ... }` blocks (x86-32 SSE, as written in the reference's src/mcemath) into C statements
with SSE intrinsics, instruction by instruction. TEST INFRASTRUCTURE (part of the shim build of the reference, oracle/ref_shim).

The point: oracle/_ref is "the reference's own arithmetic text run here". gcc cannot assemble MSVC inline asm, so every block is
rewritten 1:1 — one C statement per instruction, same order, same operands, same immediates — by this file, never by hand:

    xmm0..7            -> __m128 x_[0..7]          (live inside one block, as in the asm)
    eax..edi           -> char* r_eax ...          (pointers and small integers alike)
    mov reg, var|imm   -> r_reg = (char*)(intptr_t)(var)
    lea reg, var       -> r_reg = (char*)&(var)            lea reg, [var] likewise
    movaps / movss / addps / subps / mulps / divps / maxps / minps / andps / xorps / cmpeqps / haddps / movhlps / movlhps /
    shufps / rcpps / rsqrtss / sqrtps / sqrtss     -> the intrinsic of the same instruction (rcpps and rsqrtss stay the HARDWARE
                                                      approximations: the oracle is "the reference on this host")
    add / sub / inc / xor reg / cmp + jl / sub + jz / labels   -> integer statements, flags variables and goto
    prefetchnta                                   -> nothing

Anything unknown aborts the build: never guess.
"""
import re

ASM_BLOCK = re.compile(r"__asm\s*\{(.*?)\n[ \t]*\}", re.S)
GPR = r"e[a-d]x|esi|edi"


class AsmError(Exception):
    pass


def _imm(tok):
    tok = tok.strip()
    m = re.match(r"^([0-9][0-9a-fA-F]*)h$", tok)
    if m:
        return str(int(m.group(1), 16))
    try:
        return str(int(tok, 0))
    except ValueError:
        raise AsmError("not an immediate: %r" % tok)


def translate_block(inner, what):
    lines = []
    for raw in inner.split("\n"):
        code = raw.split(";")[0].split("//")[0].strip()
        if code:
            lines.append(code)
    out = ["{ /* shim: MSVC asm block translated instruction by instruction (%s) */" % what,
           "\t__m128 x_[8]; char *r_eax = 0, *r_ebx = 0, *r_ecx = 0, *r_edx = 0, *r_esi = 0, *r_edi = 0; intptr_t cmp_a_ = 0, cmp_b_ = 0; int zf_ = 0;",
           "\t(void)x_; (void)r_eax; (void)r_ebx; (void)r_ecx; (void)r_edx; (void)r_esi; (void)r_edi; (void)cmp_a_; (void)cmp_b_; (void)zf_;"]

    def xmm(tok):
        m = re.match(r"^xmm([0-7])$", tok.strip())
        if not m:
            raise AsmError("%s: expected an xmm register, got %r" % (what, tok))
        return "x_[%s]" % m.group(1)

    def is_xmm(tok):
        return re.match(r"^xmm[0-7]$", tok.strip()) is not None

    def is_gpr(tok):
        return re.match(r"^(%s)$" % GPR, tok.strip()) is not None

    def mem(tok):
        """[reg], [reg + off], [reg+off], [var] -> an lvalue-less address expression (char*)"""
        tok = tok.strip()
        if not (tok.startswith("[") and tok.endswith("]")):
            raise AsmError("%s: expected a memory operand, got %r" % (what, tok))
        e = tok[1:-1].strip()
        m = re.match(r"^(%s)\s*(?:\+\s*(\w+))?$" % GPR, e)
        if m:
            return "(r_%s + %s)" % (m.group(1), _imm(m.group(2)) if m.group(2) else "0")
        if re.match(r"^[A-Za-z_]\w*$", e):
            return "((char*)&(%s))" % e
        raise AsmError("%s: cannot translate address %r" % (what, tok))

    def src128(tok):
        return xmm(tok) if is_xmm(tok) else "_mm_load_ps((const float*)%s)" % mem(tok)

    binops = {"addps": "_mm_add_ps", "subps": "_mm_sub_ps", "mulps": "_mm_mul_ps", "divps": "_mm_div_ps", "maxps": "_mm_max_ps",
              "minps": "_mm_min_ps", "andps": "_mm_and_ps", "xorps": "_mm_xor_ps", "cmpeqps": "_mm_cmpeq_ps", "haddps": "_mm_hadd_ps",
              "movhlps": "_mm_movehl_ps", "movlhps": "_mm_movelh_ps"}
    for code in lines:
        m = re.match(r"^([A-Za-z_]\w*)\s*:\s*;?$", code)
        if m:                                               # a label
            out.append("%s:;" % m.group(1))
            continue
        m = re.match(r"^(\w+)\s*(.*)$", code)
        if not m:
            raise AsmError("%s: cannot parse %r" % (what, code))
        op = m.group(1).lower()
        args = [a.strip() for a in m.group(2).split(",")] if m.group(2).strip() else []
        if op == "mov" and len(args) == 2 and is_gpr(args[0]):
            if is_gpr(args[1]):
                out.append("\tr_%s = r_%s;" % (args[0], args[1]))
            elif re.match(r"^[A-Za-z_]\w*$", args[1]):
                out.append("\tr_%s = (char*)(intptr_t)(%s);" % (args[0], args[1]))
            else:
                out.append("\tr_%s = (char*)(intptr_t)%s;" % (args[0], _imm(args[1])))
        elif op == "lea" and len(args) == 2 and is_gpr(args[0]):
            a = args[1]
            out.append("\tr_%s = %s;" % (args[0], mem(a) if a.startswith("[") else "((char*)&(%s))" % a))
        elif op == "movaps" and len(args) == 2:
            if is_xmm(args[0]):
                out.append("\t%s = %s;" % (xmm(args[0]), src128(args[1])))
            elif is_xmm(args[1]):
                out.append("\t_mm_store_ps((float*)%s, %s);" % (mem(args[0]), xmm(args[1])))
            else:
                raise AsmError("%s: movaps %r" % (what, code))
        elif op == "movss" and len(args) == 2 and is_xmm(args[0]):
            if is_xmm(args[1]):
                out.append("\t%s = _mm_move_ss(%s, %s);" % (xmm(args[0]), xmm(args[0]), xmm(args[1])))
            elif args[1].startswith("["):
                out.append("\t%s = _mm_load_ss((const float*)%s);" % (xmm(args[0]), mem(args[1])))
            else:
                out.append("\t%s = _mm_load_ss(&(%s));" % (xmm(args[0]), args[1]))
        elif op in binops and len(args) == 2:
            out.append("\t%s = %s(%s, %s);" % (xmm(args[0]), binops[op], xmm(args[0]), src128(args[1])))
        elif op == "shufps" and len(args) == 3:
            out.append("\t%s = _mm_shuffle_ps(%s, %s, %s);" % (xmm(args[0]), xmm(args[0]), src128(args[1]), _imm(args[2])))
        elif op == "rcpps" and len(args) == 2:
            out.append("\t%s = _mm_rcp_ps(%s);" % (xmm(args[0]), src128(args[1])))
        elif op == "sqrtps" and len(args) == 2:
            out.append("\t%s = _mm_sqrt_ps(%s);" % (xmm(args[0]), src128(args[1])))
        elif op in ("rsqrtss", "sqrtss") and len(args) == 2 and is_xmm(args[1]):
            fn = "_mm_rsqrt_ss" if op == "rsqrtss" else "_mm_sqrt_ss"
            # only lane 0 of the destination changes
            out.append("\t%s = _mm_move_ss(%s, %s(%s));" % (xmm(args[0]), xmm(args[0]), fn, xmm(args[1])))
        elif op in ("add", "sub") and len(args) == 2 and is_gpr(args[0]):
            rhs = ("(intptr_t)r_%s" % args[1]) if is_gpr(args[1]) else (("(intptr_t)(%s)" % args[1]) if re.match(r"^[A-Za-z_]\w*$", args[1]) else _imm(args[1]))
            out.append("\tr_%s %s= %s; zf_ = (0 == r_%s);" % (args[0], "+" if op == "add" else "-", rhs, args[0]))
        elif op == "inc" and len(args) == 1 and is_gpr(args[0]):
            out.append("\tr_%s += 1; zf_ = (0 == r_%s);" % (args[0], args[0]))
        elif op == "xor" and len(args) == 2 and is_gpr(args[0]) and args[0] == args[1]:
            out.append("\tr_%s = 0; zf_ = 1;" % args[0])
        elif op == "cmp" and len(args) == 2 and is_gpr(args[0]) and is_gpr(args[1]):
            out.append("\tcmp_a_ = (intptr_t)r_%s; cmp_b_ = (intptr_t)r_%s; zf_ = (cmp_a_ == cmp_b_);" % (args[0], args[1]))
        elif op == "jl" and len(args) == 1:
            out.append("\tif(cmp_a_ < cmp_b_) goto %s;" % args[0])
        elif op == "jz" and len(args) == 1:
            out.append("\tif(zf_) goto %s;" % args[0])
        elif op == "prefetchnta":
            out.append("\t/* prefetchnta */")
        else:
            raise AsmError("%s: unsupported asm statement %r" % (what, code))
    out.append("}")
    return "\n".join(out)


def translate_source(text, what):
    """Every __asm block of `text` translated in place. Returns (text, number of blocks)."""
    count = [0]

    def repl(m):
        count[0] += 1
        return translate_block(m.group(1), "%s, block %d" % (what, count[0]))
    return ASM_BLOCK.sub(repl, text), count[0]
