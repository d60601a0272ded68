"""CPU: the measurement tools at least parse (they only run on a GPU box), and the shell helpers
name files that exist."""
import py_compile
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_python_tools_compile(tmp_path):
    tools = sorted((ROOT / "tools").glob("*.py")) + sorted((ROOT / "tools" / "microbench").glob("*.py"))
    assert len(tools) >= 8
    for t in tools + [ROOT / "bench.py", ROOT / "__graft_entry__.py"]:
        py_compile.compile(str(t), cfile=str(tmp_path / (t.stem + ".pyc")), doraise=True)


def test_shell_tools_reference_existing_files():
    for sh in sorted((ROOT / "tools").glob("*.sh")):
        for rel in re.findall(r"(?:python|bash) ((?:tools/|tests/|bench\.py)[\w/.]*)", sh.read_text()):
            assert (ROOT / rel).exists(), f"{sh.name} runs {rel}, which does not exist"
