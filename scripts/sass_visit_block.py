"""Static opcode histogram of the node-visit block of k_trace<false,0,7> in a cubin / .so (cuobjdump -sass): the
straight-line block around the 48 I2F.U8 of one visit.  Usage: sass_visit_block.py file.cubin"""
import subprocess, sys, re, collections
cubin = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n",1)[0]
    if "k_trace" not in name or "Lb0ELi0ELi7" not in name: continue
    ops = []
    for l in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m: ops.append(m.group(2).split('.')[0])
    idx = [i for i,o in enumerate(ops) if o == "I2F"]
    # densest span holding 48 consecutive I2Fs
    best = min(range(len(idx)-47), key=lambda k: idx[k+47]-idx[k])
    a, b = idx[best], idx[best+47]
    # grow to the enclosing straight-line block: back to previous BRA/BSYNC, forward to next BRA
    s = a
    while s > 0 and ops[s-1] not in ("BRA", "BSYNC", "EXIT"): s -= 1
    e = b
    while e < len(ops)-1 and ops[e] not in ("BRA", "EXIT"): e += 1
    w = ops[s:e+1]
    c = collections.Counter(w)
    print("total", len(ops), "block", len(w), "I2F span", b-a+1, sorted(c.items(), key=lambda kv:-kv[1]))
