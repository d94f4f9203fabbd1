"""Per-kernel SASS fingerprints of the built library: lets a change that is supposed to ADD an opt-in kernel variant prove
that every kernel already validated on the GPU is bitwise unchanged (no GPU needed).

    python tools/sass_fingerprint.py --write profiles/r01_validated_sass.json     # after a GPU-validated build
    python tools/sass_fingerprint.py --check profiles/r01_validated_sass.json     # later: lists changed / new / removed kernels
"""
import hashlib, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stgcn_b200", "lib", "libstgcn_b200.so")


def fingerprints(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fp, cur, h = {}, None, None
    for line in out.splitlines():
        if "Function :" in line:
            if cur:
                fp[cur] = h.hexdigest()
            cur, h = line.split("Function :")[1].strip(), hashlib.sha256()
        elif cur and re.match(r"^\s+/\*[0-9a-f]{4}\*/", line):
            # instruction text only (drop the address and the encoding comment)
            h.update(re.sub(r"/\*[0-9a-fx ]+\*/", "", line).strip().encode())
    if cur:
        fp[cur] = h.hexdigest()
    return fp


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    fp = fingerprints(sys.argv[3] if len(sys.argv) > 3 else LIB)
    if mode == "--write":
        json.dump(fp, open(path, "w"), indent=0, sort_keys=True)
        print(f"{len(fp)} kernels -> {path}")
    else:
        ref = {k: v for k, v in json.load(open(path)).items() if k not in ("_note", "_unvalidated_opt_in")}
        changed = sorted(k for k in ref if k in fp and fp[k] != ref[k])
        removed = sorted(k for k in ref if k not in fp)
        new = sorted(k for k in fp if k not in ref)
        print(f"{len(ref)} reference kernels: {len(changed)} changed, {len(removed)} removed (or renamed), {len(new)} new")
        for tag, lst in (("changed", changed), ("removed", removed), ("new", new)):
            for k in lst:
                print(f"  {tag}: {k[:140]}")
        sys.exit(1 if changed else 0)
