"""Documentation hygiene (CPU): every profiles/ artefact and test file that DESIGN.md, README.md, INTEGRATION.md or
profiles/README.md name exists in the tree, so the evidence the documents cite can actually be opened."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")]


def _expand(name):
    m = re.search(r"\{([^}]*)\}", name)               # r01_bench_fe_{4,8}gpu.json
    if not m:
        return [name]
    return [name[:m.start()] + alt + name[m.end():] for alt in m.group(1).split(",")]


def test_cited_profiles_and_tests_exist():
    missing = []
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for tok in re.findall(r"`([^`\s]+)`", text):
            tok = tok.rstrip(".,;:")
            base = os.path.basename(tok)
            if re.fullmatch(r"r\d\d_[\w{},.*-]+\.(json|csv|txt)", base):
                for name in _expand(base):
                    if not glob.glob(os.path.join(ROOT, "profiles", name)):
                        missing.append((doc, name))
            elif re.fullmatch(r"test_\w+\.py", base):
                if not os.path.exists(os.path.join(ROOT, "tests", base)):
                    missing.append((doc, base))
            elif tok.startswith(("scripts/", "profiles/microbench/")) and tok.endswith((".py", ".cu", ".sh")):
                if not os.path.exists(os.path.join(ROOT, tok)):
                    missing.append((doc, tok))
    assert not missing, missing
