"""Writes tests/golden/voc_ref.npz: synthetic vocabularies (seeded, tests/voc_cases.py), descriptors and the BowVector / FeatureVector that the
reference's OWN vendored DBoW2 (oracle/_ref/libref_voc.so: TemplatedVocabulary::loadFromTextFile + transform, compiled unmodified) returns for them.

    python tests/golden/make_voc_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle
import voc_cases as vc


def main():
    R = oracle.ref_voc()
    if R is None:
        raise SystemExit("oracle/_ref/libref_voc.so not built (needs /root/reference): make -C oracle ref")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for j, (k, L, irregular, levelsup) in enumerate(vc.CASES):
            rng = np.random.default_rng(500 + j)
            tree = vc.make_tree(rng, k, L, irregular)
            feats = vc.features(rng, tree)
            path = os.path.join(tmp, "voc%d.txt" % j)
            vc.write_voc_text(path, k, L, tree)
            ans = vc.vectors_ref(R, path, feats, levelsup)
            for name, v in ans.items():
                out["c%d.%s" % (j, name)] = v
            print(j, (k, L, irregular, levelsup), "nodes", len(tree[0]), "words", int(ans["words"]), "bow", len(ans["bow_words"]), "fv nodes", len(ans["fv_nodes"]))
    path = os.path.join(HERE, "voc_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
