"""ORACLE / TEST INFRASTRUCTURE ONLY — recipe that vendors the reference's own hot-path modules for the CPU baseline.

Copies, unmodified, the few files of /root/reference/Modules that define the path (cells, make_mlp, the two models and
their YAML configs) into oracle/_ref/Modules/. oracle/_ref/ is git-ignored (no reference source enters the history) but
is NOT gpurun-ignored, so the copy travels to the GPU box, where `bench.py --impl reference` and the `cpu_baseline` leg
import it through oracle/reference_harness.py (third-party CUDA packages restated by oracle/stubs) and time the
reference's own code on the host cores (`kind: "reference"`). Run by __graft_entry__.build() wherever /root/reference
exists; a no-op elsewhere.
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("HGNN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = [
    "Modules/gnn_utils.py",
    "Modules/utils.py",
    "Modules/tracking_utils.py",
    "Modules/EdgeClassifier/edge_classifier_base.py",
    "Modules/EdgeClassifier/Models/IN.py",
    "Modules/EdgeClassifier/Configs/IN.yaml",
    "Modules/BipartiteClassification/bipartite_classification_base.py",
    "Modules/BipartiteClassification/Models/HGNN_GMM.py",
    "Modules/BipartiteClassification/Configs/HGNN_GMM.yaml",
]


def main():
    if not os.path.isdir(os.path.join(SRC, "Modules")):
        print("make_ref: no reference tree at", SRC, "- nothing to do")
        return False
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    print("make_ref: vendored", len(FILES), "reference files into", DST)
    return True


if __name__ == "__main__":
    main()
