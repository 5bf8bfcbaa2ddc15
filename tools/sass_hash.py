"""Per-kernel SASS fingerprint of a built library: `python tools/sass_hash.py lib.so > a.txt`, edit, rebuild, diff.
Used to prove that a gated experiment or a new template instantiation leaves the validated kernels bit-identical."""
import hashlib, re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True, check=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name, body = part.split("\n", 1)
    ins = [re.sub(r"/\*[0-9a-f]+\*/", "", l).strip() for l in body.splitlines() if re.match(r"\s*/\*[0-9a-f]{4,5}\*/", l)]
    print(hashlib.md5("\n".join(ins).encode()).hexdigest()[:12], len(ins), name.strip())
