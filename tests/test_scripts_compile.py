"""The measurement scripts are part of the evidence trail (profiles/README.md names the command behind every file):
keep them syntactically alive in the CPU suite. Nothing here touches a GPU."""

import pathlib
import py_compile
import subprocess

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_python_scripts_compile():
    scripts = sorted((ROOT / "scripts").glob("*.py")) + [ROOT / "bench.py", ROOT / "__graft_entry__.py"]
    assert len(scripts) >= 6
    for path in scripts:
        py_compile.compile(str(path), doraise=True)


def test_round_check_script_parses():
    res = subprocess.run(["bash", "-n", str(ROOT / "scripts" / "gpu_round_check.sh")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_every_profile_the_docs_cite_exists():
    import re

    cited = set()
    for doc in ("DESIGN.md", "README.md", "profiles/README.md"):
        cited |= set(re.findall(r"`(?:profiles/)?(r1[a-e]?_[A-Za-z0-9_.]+\.(?:json|jsonl|txt|csv|log))`", (ROOT / doc).read_text()))
    assert len(cited) >= 15
    missing = sorted(name for name in cited if not (ROOT / "profiles" / name).exists())
    assert not missing, missing
