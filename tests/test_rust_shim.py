"""The Rust side of the drop-in (rust/gpu.rs + rust/patches/) cannot be compiled here — no cargo / rustc in the image — so
it is held to what CAN be checked mechanically:

* every `#[repr(C)]` struct and every `extern "C"` prototype of gpu.rs is the transcription of include/rayrs_b200.h:
  same fields in the same order with the same scalar types / array lengths / pointer-ness, same constants;
* the patches are purely additive (main.rs too: the call goes in front of the rayon tile loop, which stays, compiled out)
  and, where the reference sources are present (this container; not the GPU box), apply cleanly to them;
* everything gpu.rs uses from the patched modules (`BvhCursor`, `FlatGeom`, `FlatMaterial`, `Hittable::flatten`, ...) is
  defined by the patches.
"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "rayrs_b200.h").read_text()
GPU_RS = (ROOT / "rust" / "gpu.rs").read_text()
PATCHES = sorted((ROOT / "rust" / "patches").glob("*.patch"))
REFERENCE = Path("/root/reference")

C_SCALAR = {"uint32_t": "u32", "int32_t": "i32", "uint64_t": "u64", "int64_t": "i64", "double": "f64", "float": "f32",
            "int": "c_int", "char": "c_char", "uint8_t": "u8", "size_t": "usize", "void": "c_void"}


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def _c_structs():
    """name -> [(field, rust-style type string)] for every `typedef struct X { ... } X;` of the header"""
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", _strip_c_comments(HEADER), flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            is_ptr = "*" in decl
            decl_np = decl.replace("*", " ").replace("const ", "")
            parts = decl_np.split()
            ctype, names = parts[0], " ".join(parts[1:])
            for name in names.split(","):
                name = name.strip()
                arr = re.match(r"(\w+)\[(\d+)\]$", name)
                base = C_SCALAR.get(ctype, ctype)
                if is_ptr:
                    fields.append((name, "ptr " + base))
                elif arr:
                    fields.append((arr.group(1), f"[{base}; {arr.group(2)}]"))
                else:
                    fields.append((name, base))
        out[m.group(1)] = fields
    return out


def _rust_structs():
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[derive\([^\]]*\)\]\s*)?pub struct (\w+) \{(.*?)\n\}", GPU_RS, flags=re.S):
        fields = []
        for line in m.group(2).splitlines():
            line = line.split("//")[0].strip().rstrip(",")
            fm = re.match(r"(?:pub )?(\w+): (.+)$", line)
            if not fm:
                continue
            name, typ = fm.group(1), fm.group(2).strip()
            if typ.startswith("*const ") or typ.startswith("*mut "):
                typ = "ptr " + typ.split(" ", 1)[1]
            fields.append((name.rstrip("_"), typ))
        out[m.group(1)] = fields
    return out


def test_repr_c_structs_are_the_headers():
    c, r = _c_structs(), _rust_structs()
    checked = 0
    for name, rf in r.items():
        if name in ("RrsScene", "RrsComm"):  # opaque handles: `typedef struct X X;` in the header
            assert re.search(rf"typedef struct {name} {name};", HEADER)
            continue
        assert name in c, f"gpu.rs declares {name}, the header does not"
        assert rf == c[name], f"{name}: gpu.rs {rf} != header {c[name]}"
        checked += 1
    # the structs that cross the FFI on the render path must all be there
    for need in ("RrsPrim", "RrsMaterial", "RrsEmission", "RrsNode", "RrsNodeF64", "RrsSceneDesc", "RrsCamera", "RrsRenderParams",
                 "RrsStats"):
        assert need in r, need
    assert checked >= 9


def _c_prototypes():
    out = {}
    text = _strip_c_comments(HEADER)
    for m in re.finditer(r"^(int|void|const char\*) (rrs_\w+)\((.*?)\);", text, flags=re.S | re.M):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        params = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                stars = a.count("*")
                toks = a.replace("*", " ").replace("const ", "").split()
                base = C_SCALAR.get(toks[0], toks[0])
                params.append("ptr" * 0 + ("*" * stars) + base)
        out[name] = ({"int": "c_int", "void": "", "const char*": "*c_char"}[ret], params)
    return out


def _rust_prototypes():
    out = {}
    block = re.search(r'extern "C" \{(.*?)\n\}', GPU_RS, flags=re.S).group(1)
    for m in re.finditer(r"fn (rrs_\w+)\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, args, ret = m.group(1), " ".join(m.group(2).split()), (m.group(3) or "").strip()
        params = []
        for a in filter(None, (x.strip() for x in args.split(","))):
            typ = a.split(":", 1)[1].strip()
            stars = typ.count("*")
            base = typ.replace("*const", "").replace("*mut", "").strip()
            params.append("*" * stars + base)
        ret = ret.replace("*const ", "*").replace("*mut ", "*")
        out[name] = (ret, params)
    return out


def test_extern_block_matches_the_header_prototypes():
    c, r = _c_prototypes(), _rust_prototypes()
    assert len(r) >= 10
    for name, (ret, params) in r.items():
        assert name in c, f"gpu.rs binds {name}, which the header does not declare"
        assert ret == c[name][0], (name, ret, c[name][0])
        assert params == c[name][1], (name, params, c[name][1])
    # the render path: scene creation (one and several GPUs), both render calls, the communicator, stats, errors
    for need in ("rrs_scene_create", "rrs_scene_create_multi", "rrs_scene_destroy", "rrs_render", "rrs_render_multi",
                 "rrs_comm_init_all", "rrs_comm_destroy", "rrs_stats", "rrs_last_error", "rrs_device_count"):
        assert need in r, need


def test_constants_match_the_header():
    def c_define(name):
        m = re.search(rf"#define {name} (0x[0-9A-Fa-f]+|\d+)u?", HEADER)
        assert m, name
        return int(m.group(1), 0)

    def rust_const(name):
        m = re.search(rf"pub const {name}: u32 = ([0-9A-Fa-fx_]+);", GPU_RS)
        assert m, name
        return int(m.group(1).replace("_", ""), 0)

    for name in ("RRS_ABI_VERSION", "RRS_REF_LEAF", "RRS_REF_EMPTY"):
        assert rust_const(name) == c_define(name), name
    # material tags travel as the declaration order of `enum Material` (material.rs:57-68) == RrsMaterialTag
    tags = re.search(r"typedef enum RrsMaterialTag \{(.*?)\}", HEADER, flags=re.S).group(1)
    order = [t.split("=")[0].strip() for t in tags.split(",") if t.strip()]
    assert [int(t.split("=")[1]) for t in tags.split(",") if t.strip()] == list(range(len(order)))
    assert order[0] == "RRS_MAT_LAMBERTIAN" and order[-1] == "RRS_MAT_NO_REFLECT" and len(order) == 9


def test_patches_only_add_lines():
    assert len(PATCHES) == 5
    for p in PATCHES:
        removed = [l for l in p.read_text().splitlines() if l.startswith("-") and not l.startswith("---")]
        assert not removed, (p.name, removed[:3])  # main.rs.patch too: the CPU tile loop stays, compiled out


def test_gpu_rs_uses_only_what_the_patches_define():
    added = "\n".join(l[1:] for p in PATCHES for l in p.read_text().splitlines() if l.startswith("+") and not l.startswith("+++"))
    for item in ("BvhCursor", "FlatGeom", "FlatMaterial"):
        assert re.search(rf"(struct|enum) {item}\b", added), f"{item} is imported by gpu.rs and defined by no patch"
    assert "pub mod gpu;" in added
    for method in ("fn cursor", "fn flatten", "fn flat"):
        assert method in added, method
    # every method gpu.rs calls on a cursor exists on BvhCursor
    cursor_impl = added[added.index("impl<'a> BvhCursor<'a>"):]
    defined = set(re.findall(r"pub\(crate\) fn (\w+)", cursor_impl))
    used = set(re.findall(r"\b(?:cursor|node|child|c|root)\.(\w+)\(", GPU_RS)) & {"bbox", "children", "object", "is_leaf", "len"}
    assert used <= defined | {"len"}, (used, defined)


@pytest.mark.skipif(not REFERENCE.exists() or shutil.which("patch") is None, reason="reference sources / patch(1) not present")
def test_patches_apply_to_the_reference(tmp_path):
    """on a scratch copy: /root/reference is read-only and nothing of it enters the repository"""
    for sub in ("rayrs-lib/src", "rayrs/src"):
        (tmp_path / sub).mkdir(parents=True)
        for f in (REFERENCE / sub).glob("*.rs"):
            shutil.copy(f, tmp_path / sub / f.name)
    for p in PATCHES:
        r = subprocess.run(["patch", "-p1", "--forward", "--batch", "-i", str(p)], cwd=tmp_path, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, (p.name, r.stdout)
        assert "FAILED" not in r.stdout and "fuzz" not in r.stdout, (p.name, r.stdout)
    lib = (tmp_path / "rayrs-lib/src/lib.rs").read_text()
    assert "pub mod gpu;" in lib
    main = (tmp_path / "rayrs/src/main.rs").read_text()
    # the call sits before the tile loop, which is compiled out (`#[cfg(any())]` is never true)
    assert main.index("render_gpu(") < main.index("#[cfg(any())]") < main.index("into_par_iter")
    # the patched sources still balance their braces (a cheap stand-in for the compiler that is not here)
    for f in ("rayrs-lib/src/bvh.rs", "rayrs-lib/src/geometry.rs", "rayrs-lib/src/material.rs", "rayrs/src/main.rs"):
        text = re.sub(r"//.*", "", (tmp_path / f).read_text())
        text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text)
        text = re.sub(r"'(?:\\.|[^'\\])'", "' '", text)
        assert text.count("{") == text.count("}"), f
        assert text.count("(") == text.count(")"), f


def test_gpu_rs_balances_its_delimiters():
    text = re.sub(r"//.*", "", GPU_RS)
    text = re.sub(r'"(?:\\.|[^"\\])*"', '""', text)
    text = re.sub(r"'(?:\\.|[^'\\])'", "' '", text)
    for a, b in ("{}", "()", "[]"):
        assert text.count(a) == text.count(b), (a, text.count(a), text.count(b))
