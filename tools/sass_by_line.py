"""SASS instruction count per source line of one kernel: nvdisasm -g <cubin> > all.sass, then
python tools/sass_by_line.py all.sass <mangled-name-substring> [top]"""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur = None
inside = False
cnt = collections.Counter()
for ln in open(path):
    if ln.startswith(".text."):
        inside = key in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln) and cur:
        cnt[cur] += 1
print("total", sum(cnt.values()))
for (f, l), c in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    print(f"{f}:{l}  {c}")
