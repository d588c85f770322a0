"""Instruction and local-memory (spill) counts per source line of one kernel, from a -lineinfo object file.
   python tools/sass_lines.py <object.o> <kernel-name-substring> [top]
Needs cuobjdump + nvdisasm (no GPU)."""
import collections, pathlib, re, subprocess, sys, tempfile

obj, name = pathlib.Path(sys.argv[1]).resolve(), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", str(obj)], cwd=d, check=True, capture_output=True)
    cubin = next(pathlib.Path(d).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
inside, cur = False, None
tot, loc = collections.Counter(), collections.Counter()
for line in dis.splitlines():
    if line.startswith(".text."):
        inside = name in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        tot[cur] += 1
        if m.group(2).startswith(("LDL", "STL")):
            loc[cur] += 1
print(f"{name}: {sum(tot.values())} instructions, {sum(loc.values())} local-memory")
for k, v in tot.most_common(top):
    print(f"  {k[0]}:{k[1]:<6d} {v:6d}  local {loc[k]}")
