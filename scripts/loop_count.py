"""Compile csrc/bellman_window.cu (or a given file) and print the instruction mix of the inner
loops of the window-kernel instantiations: python scripts/loop_count.py [file.cu] [name-filter]"""
import re, subprocess, sys, os, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "optimal-control-dynamic-programming_b200/csrc/bellman_window.cu")
flt = sys.argv[2] if len(sys.argv) > 2 else "Lb1ELb1E"
csrc = os.path.join(ROOT, "optimal-control-dynamic-programming_b200/csrc")
with tempfile.TemporaryDirectory() as td:
    cub = os.path.join(td, "w.cubin")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-I" + csrc,
                    "-I" + os.path.join(ROOT, "include"), "-cubin", "-o", cub, src], check=True)
    sass = subprocess.run(["cuobjdump", "-sass", cub], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
for f in funcs:
    name = f.split("\n")[0]
    if "k_stage_window" not in name or flt not in name:
        continue
    lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    addr = {int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1), 16): i for i, l in enumerate(lines)}
    print(name[:110], len(lines), "instr")
    for i, l in enumerate(lines):
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            if j < i and 100 < i - j < 600:
                ops = {}
                for b in lines[j:i + 1]:
                    t = re.sub(r"^\s+/\*[0-9a-f]{4}\*/\s+", "", b); t = re.sub(r"^@!?U?P\d\s+", "", t)
                    op = t.split()[0].split(".")[0] if t.split() else ""
                    ops[op] = ops.get(op, 0) + 1
                if ops.get("LDS", 0) >= 8:
                    print("   loop %4d instr:" % (i - j + 1), sorted(ops.items(), key=lambda kv: -kv[1])[:12])
