"""The cgo shim under integration/go cannot be compiled in this image (no Go toolchain).  These checks keep its C side from drifting: every
C.kr_* name it uses exists in include/kr_engine.h, every field of a C struct literal is a field of that struct, every KR_* constant is
defined, and the generated column file matches its generator."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "kr_engine.h")).read()
GO_FILES = sorted(glob.glob(os.path.join(ROOT, "integration", "go", "krengine", "*.go")))


def struct_fields(name: str) -> set[str]:
    m = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", HEADER, re.S)
    assert m, f"struct {name} is not in the header"
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = set()
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            ident = re.findall(r"[A-Za-z_][A-Za-z_0-9]*", re.sub(r"\[.*?\]", "", part))
            if ident:
                fields.add(ident[-1])
    return fields


def test_the_shim_is_there():
    assert {os.path.basename(f) for f in GO_FILES} >= {"doc.go", "engine.go", "columns_gen.go", "packer.go", "podbuild.go", "batcher.go"}


def test_every_c_name_the_shim_uses_is_in_the_header():
    for path in GO_FILES:
        src = open(path).read()
        for name in set(re.findall(r"\bC\.(kr_[a-z_0-9]+)", src)):
            assert re.search(r"\b" + name + r"\b", HEADER), f"{os.path.basename(path)}: C.{name} is not declared in kr_engine.h"
        for const in set(re.findall(r"\bC\.(KR_[A-Z_0-9]+)", src)):
            assert re.search(r"\b" + const + r"\b", HEADER), f"{os.path.basename(path)}: C.{const} is not defined in kr_engine.h"


def test_struct_literals_name_real_fields():
    seen = 0
    for path in GO_FILES:
        src = open(path).read()
        for m in re.finditer(r"\bC\.(kr_[a-z_0-9]+)\{", src):
            # the literal's body up to the matching brace
            depth, i = 1, m.end()
            while depth and i < len(src):
                depth += {"{": 1, "}": -1}.get(src[i], 0)
                i += 1
            body = src[m.end():i - 1]
            top, depth = "", 0
            for ch in body:  # only the literal's own keys, not those of nested literals / calls
                depth += {"{": 1, "(": 1, "}": -1, ")": -1}.get(ch, 0)
                top += ch if depth == 0 else " "
            keys = re.findall(r"(?:^|,)\s*([a-z_][a-z_0-9]*)\s*:", top)
            if keys:
                fields = struct_fields(m.group(1))
                for k in keys:
                    assert k in fields, f"{os.path.basename(path)}: {m.group(1)} has no field {k}"
                    seen += 1
    assert seen > 60
    # field reads / writes through a variable of a known struct type: x.field where the file declares `var x C.T` or `x := C.T{`
    checked = 0
    for path in GO_FILES:
        for src in re.split(r"\nfunc ", open(path).read())[1:]:      # one function at a time: short variable names are reused across functions
            for var, typ in re.findall(r"\b([a-z][A-Za-z0-9]*)\s*:=\s*C\.(kr_[a-z_0-9]+)\{", src) + re.findall(r"\bvar\s+([a-z][A-Za-z0-9]*)\s+C\.(kr_[a-z_0-9]+)\b", src) + \
                    re.findall(r"[(,]\s*([a-z][A-Za-z0-9]*)\s+\*C\.(kr_[a-z_0-9]+)\b", src.split("{", 1)[0]):
                if not re.search(r"typedef struct " + typ + r"\s*\{", HEADER):
                    continue
                fields = struct_fields(typ)
                for f in set(re.findall(r"(?<![\w.])" + var + r"\.([a-z_][a-z_0-9]*)\b", src)):
                    assert f in fields, f"{os.path.basename(path)}: {var}.{f} is not a field of {typ}"
                    checked += 1
    assert checked > 15


def test_generated_columns_are_current(tmp_path):
    path = os.path.join(ROOT, "integration", "go", "krengine", "columns_gen.go")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_go_shim.py")], stdout=subprocess.DEVNULL)
    assert open(path).read() == before, "run tools/gen_go_shim.py and commit the result"
    # every column of kr_snapshot_bufs is there
    for col in struct_fields("kr_snapshot_bufs"):
        assert f"b.{col})" in before, col
