"""oracle/make_ref.py -- TEST INFRASTRUCTURE ONLY: stages the UNMODIFIED reference under oracle/_ref/.

The reference is pure Python (six .py files + three CSV tables), so "building" it is a copy.  The files are
copied byte for byte from /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored: like a
compiled reference .so it travels to the GPU box, where /root/reference does not exist).  Nothing is
edited; tests/test_reference_mains_gpu.py and bench.py's --impl reference arm import them from there with
the two shims of SURVEY.md Appendix F (stub ``imagenet`` module, seeded ``load_state_dict_from_url``).

    python oracle/make_ref.py            # no-op (exit 0) when /root/reference is absent
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("SLQ_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["functions.py", "resnet.py", "imagenet.py", "resnet18_main.py", "resnet34_main.py", "resnet50_main.py",
         "dataset/resnet18_deltaloss.csv", "dataset/resnet34_deltaloss.csv", "dataset/resnet50_deltaloss.csv"]


def staged():
    return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)


def make(verbose=False):
    if not os.path.isfile(os.path.join(REF_DIR, "functions.py")):
        return staged()
    os.makedirs(os.path.join(DST, "dataset"), exist_ok=True)
    digest = hashlib.sha256()
    for f in FILES:
        src, dst = os.path.join(REF_DIR, f), os.path.join(DST, f)
        data = open(src, "rb").read()
        digest.update(f.encode() + b"\0" + data)
        if not os.path.isfile(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", f)
    with open(os.path.join(DST, "SHA256"), "w") as fh:
        fh.write(digest.hexdigest() + "\n")
    return True


if __name__ == "__main__":
    ok = make(verbose=True)
    print("oracle/_ref %s" % ("ready" if ok else "not staged (no reference tree here)"))
    sys.exit(0)
