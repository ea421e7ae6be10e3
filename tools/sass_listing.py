"""SASS evidence of the shipped libhmz.so (no GPU needed): full listings of the tree and env kernels as text, of the two big
tensor-core kernels gzipped, and one summary (opcode histograms + the lines that prove tcgen05 / TMEM / TMA / 256-bit
accesses).      python tools/sass_listing.py [round-prefix, default r02]"""
import collections, gzip, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT = os.path.join(ROOT, "profiles", "sass")
os.makedirs(OUT, exist_ok=True)
text = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "muzero-hanoi_b200", "libhmz.so")], check=True,
                      capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in text.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    if cur:
        funcs[cur].append(line)


def find(key):
    return next(k for k in funcs if key in k)


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


plain = {"tree_search_backup_select": "search_backup_selectILb0ELb1", "tree_search_select": "search_selectE",
         "env_step_vec4": "env_step_vec4ILb0ELi1", "env_step_random_vec4": "env_step_random_vec4ILi1"}
zipped = {"net_tc_recurrent": "net_tcILb0", "search_persistent": "search_persistent", "net_x3_recurrent": "net_x3ILb0"}
for name, key in plain.items():
    with open(os.path.join(OUT, f"{R}_{name}.sass"), "w") as f:
        f.write("\n".join(funcs[find(key)]) + "\n")
for name, key in zipped.items():
    with gzip.open(os.path.join(OUT, f"{R}_{name}.sass.gz"), "wt") as f:
        f.write("\n".join(funcs[find(key)]) + "\n")
proof = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "LDG.E.ENL2.256", "STG.E.ENL2.256", "LDG.E.128", "STG.E.EF.128", "CCTL.IVALL",
         "MEMBAR.ALL.GPU", "REDG.E.ADD.STRONG.GPU", "ATOMG.E.ADD.STRONG.GPU", "DFMA", "BSSY"]
with open(os.path.join(OUT, f"{R}_summary.txt"), "w") as f:
    for key in ("search_backup_selectILb0ELb1", "env_step_vec4ILb0ELi1", "env_step_random_vec4ILi1", "net_tcILb0", "search_persistent", "net_x3ILb0"):
        k = find(key)
        ins = [l for l in funcs[k] if re.match(r"\s*/\*[0-9a-f]+\*/", l)]
        ops = collections.Counter()
        for l in ins:
            toks = l.split()[1:]
            op = toks[1] if toks[0].startswith("@") else toks[0]
            ops[op.rstrip(";").split(".")[0]] += 1
        f.write(f"== {demangle(k)}\n   {len(ins)} instructions; by opcode: " + ", ".join(f"{o} {c}" for o, c in ops.most_common(30)) + "\n")
        for p in proof:
            n = sum(1 for l in ins if p in l)
            if n:
                first = next(l for l in ins if p in l).strip()
                f.write(f"   {p:24s} x{n:4d}   e.g. {first[:120]}\n")
print("\n".join(sorted(os.listdir(OUT))))
