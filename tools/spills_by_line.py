"""Where does a kernel spill?  usage: spills_by_line.py <nvdisasm -g output> <function substring...>"""
import re, collections, sys
lines = open(sys.argv[1]).read().split('\n')
idx = [i for i, l in enumerate(lines) if l.startswith('.text.')]
for i in idx:
    if all(s in lines[i] for s in sys.argv[2:]):
        start, end = i, min([j for j in idx if j > i] + [len(lines)])
        break
cur = None
cnt = collections.Counter()
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if re.search(r'\b(STL|LDL)', l):
        cnt[(cur, 'STL' if 'STL' in l else 'LDL')] += 1
print(lines[start])
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:25]:
    print(k, v)
