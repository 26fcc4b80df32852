"""Writes tests/golden/dict_ref.npz: canonical marker patches and what the reference's OWN marker identification (Thirdparty/aruco/aruco/dictionary.cpp,
dictionary_based.cpp, markerlabeler.cpp compiled unmodified into oracle/_ref/libref_dict.so, oracle/Makefile) answers for them, a digest of every
predefined dictionary's code table as dictionary.cpp holds it, and Dictionary::getMarkerImage_id renderings.  Replayed by tests/test_oracle_dict_vs_ref.py
on boxes without the reference.

    python tests/golden/make_dict_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dict_cases as dc
import oracle


def main():
    R = oracle.ref_dict()
    if R is None:
        raise SystemExit("oracle/_ref/libref_dict.so not built (needs /root/reference): make -C oracle ref")
    out = {}
    for name in dc.DICTS:
        nb, tau, codes = oracle.dictionary_codes(name, impl=R)
        out["table.%s" % name] = np.frombuffer(hashlib.sha256(codes.tobytes()).digest(), np.uint8)
        out["meta.%s" % name] = np.array([nb, tau, len(codes)], np.int32)
    for name in dc.DECODE_DICTS:
        patches = dc.patches_for(name, frames=2)
        ans = np.array([oracle.decode_patch(p, name, impl=R) for p in patches], np.int32)
        out["patches.%s" % name] = patches
        out["answers.%s" % name] = ans
        print(name, len(patches), "patches,", int(ans[:, 0].sum()), "identified")
        ids = dc.RENDER_IDS[name]
        out["render.%s" % name] = np.stack([dc.ref_marker_image(R, name, i, 4) for i in ids])
    path = os.path.join(HERE, "dict_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
