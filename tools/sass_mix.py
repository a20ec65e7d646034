"""Static instruction mix of the hot kernels of libmptg.so (cuobjdump -sass), written to profiles/ as evidence for the
instruction-count statements of DESIGN.md.   python tools/sass_mix.py [regex ...] > profiles/r2_sass_mix.txt
Counts are STATIC (instructions in the cubin, loops counted once); the dynamic counts quoted next to them come from ncu
(smsp__inst_executed, source page).  Also lists, per kernel, the TMA / bulk-copy / cp.async / packed-half / tensor mnemonics."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "mpt_b200" / "_lib" / "libmptg.so"
pats = sys.argv[1:] or ["knnSe3Kernel", "knnBvhKernel<float, *\\(int\\)0", "knnBruteL1Kernel<double", "meshFlatKernel", "flatLinkKernel<float, mptg::nao",
                        "flatLinkKernel<double, mptg::ArmValidator<double, \\(int\\)8", "validKernel<float, mptg::nao"]
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
funcs, name = {}, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = demangle(m.group(1))
        funcs[name] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        funcs[name].append(m.group(1))
GROUPS = [("FP32 FFMA", r"^FFMA"), ("FP32 FMUL", r"^FMUL"), ("FP32 FADD", r"^FADD"), ("FP32 compare/select/minmax", r"^(FSETP|FSEL|FMNMX|FSET|FCHK)"),
          ("packed half (HFMA2/HADD2/HMUL2/HSETP2/HMNMX2)", r"^H(FMA2|ADD2|MUL2|SETP2|MNMX2|SET2)"), ("FP64 (DFMA/DADD/DMUL/DSETP)", r"^D(FMA|ADD|MUL|SETP|MNMX)"),
          ("MUFU (rcp/rsq/sqrt...)", r"^MUFU"), ("conversions (F2F/I2F/F2I/F2FP/HADD2.F32)", r"^(F2F|I2F|F2I|F2FP|I2FP)"), ("integer ALU (IADD3/IMAD/LOP3/SHF/LEA/ISETP...)", r"^(IADD|IMAD|LOP3|SHF|LEA|ISETP|IABS|POPC|FLO|BREV|PRMT|SEL|MOV|SGXT|VIADD|VIMNMX|IMNMX|UIADD|ULOP|UMOV|USHF|UIMAD|ULEA|UISETP|R2UR|S2R|S2UR|CS2R|PLOP3|UPLOP|P2R|R2P)"),
          ("global/local loads+stores (LDG/STG/LDL/STL/LD/ST)", r"^(LDG|STG|LDL|STL|LD\.|ST\.|LD$|ST$)"), ("shared loads+stores (LDS/STS/LDSM)", r"^(LDS|STS|LDSM)"), ("constant loads (LDC/ULDC)", r"^U?LDC"),
          ("async copies LDGSTS (cp.async)", r"^LDGSTS"), ("bulk copies UBLKCP (cp.async.bulk)", r"^UBLKCP"), ("TMA tensor UTMALDG/UTMASTG", r"^UTMA"), ("mbarrier SYNCS", r"^SYNCS"),
          ("warp vote/shuffle/redux (VOTE/SHFL/REDUX/MATCH)", r"^(VOTE|SHFL|REDUX|MATCH)"), ("atomics (ATOM/RED/ATOMS/ATOMG)", r"^(ATOM|RED)"), ("barriers/fences (BAR/MEMBAR/WARPSYNC/ERRBAR)", r"^(BAR|MEMBAR|WARPSYNC|ERRBAR|DEPBAR)"),
          ("branches (BRA/BSSY/BSYNC/EXIT/CALL/RET/BRX)", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|BRX|BREAK|JMP|WARPSYNC)"), ("tensor core (HMMA/UTCHMMA/tcgen05)", r"^(HMMA|IMMA|UTC|QGMMA|UTCHMMA)")]
print(f"# static SASS instruction mix, {LIB.name}, sm_100a  (tools/sass_mix.py)")
for pat in pats:
    hits = [n for n in funcs if re.search(pat, n)]
    for n in hits[:3]:
        ops = funcs[n]
        c = Counter(ops)
        print(f"\n== {n[:200]}\n   {len(ops)} instructions")
        used = 0
        for label, rx in GROUPS:
            k = sum(v for op, v in c.items() if re.match(rx, op))
            if k:
                print(f"   {label:60s} {k:7d}  {100.0 * k / len(ops):5.1f} %")
        top = ", ".join(f"{op} {v}" for op, v in c.most_common(14))
        print(f"   top mnemonics: {top}")
    if not hits:
        print(f"\n== no function matches {pat}")
