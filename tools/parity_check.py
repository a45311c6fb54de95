"""Compares the byte files written by tools/parity_dump.rs (the REAL reference crates) with this repository's committed golden
digests (tests/golden/golden_named.json, produced by the CPU oracle and matched byte for byte by the CUDA prover):
    python tools/parity_check.py <out_dir> <tag>
Exit status 0 = every digest matches: the oracle's composition (Merlin framing, label order, bincode layout) is the reference's."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out, tag = sys.argv[1], sys.argv[2]
    gold = {(c["tag"], c["kind"]): c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_named.json")))["cases"]}
    bad = 0
    for kind in ("point_add", "point_mult"):
        want = gold.get((tag, kind))
        if want is None or not os.path.exists(os.path.join(out, f"{kind}_proof.bin")):
            continue
        for what, key in (("comm", "comm_sha256"), ("comm_vars_para", "comm_vars_para_sha256"), ("comm_vars_input", "comm_vars_input_sha256"),
                          ("comm_vars", "comm_vars_sha256"), ("proof", "proof_sha256")):
            b = open(os.path.join(out, f"{kind}_{what}.bin"), "rb").read()
            ok = hashlib.sha256(b).hexdigest() == want[key] and (what != "proof" or (len(b) == want["proof_len"] and b[:64].hex() == want["proof_head"]))
            print(f"{tag}/{kind}/{what}: {'match' if ok else 'MISMATCH'} ({len(b)} bytes)")
            bad += 0 if ok else 1
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
