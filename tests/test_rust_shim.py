"""The Rust side cannot be compiled here (no cargo/rustc in the image), so these CPU tests pin the things a
build would trip over: build.rs compiles exactly the Makefile's sources with the Makefile's flags and links
the same libraries, and every `extern "C"` name the shim binds is declared in include/dawn_index.h and exported by the
built library (north_star (1); replaces Cargo.toml:36-38 / search_provider.rs:32 of the reference)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _makefile_var(name):
    src = open(os.path.join(ROOT, "dawnsearch_b200", "csrc", "Makefile")).read()
    m = re.search(rf"^{name}\s*=\s*(.*)$", src, flags=re.M)
    assert m, name
    return m.group(1).split()


def _build_rs():
    return open(os.path.join(ROOT, "rust", "build.rs")).read()


def _rs_list(name):
    m = re.search(rf"const {name}: &\[&str\] = &\[(.*?)\];", _build_rs(), flags=re.S)
    assert m, name
    return re.findall(r'"([^"]+)"', m.group(1))


def test_build_rs_compiles_every_source_of_the_makefile():
    srcs = _makefile_var("SRCS")
    assert sorted(_rs_list("SRCS")) == sorted(srcs)
    on_disk = sorted(f for f in os.listdir(os.path.join(ROOT, "dawnsearch_b200", "csrc")) if f.endswith(".cu"))
    assert on_disk == sorted(srcs), "a .cu file in csrc/ is missing from the Makefile (and build.rs)"


def test_build_rs_uses_the_makefile_flags_and_link_line():
    flags = [f for f in _makefile_var("NVCCFLAGS") if f not in ("$(ARCH)", "$(EXTRA)")]  # EXTRA: empty unless an A/B build asks
    arch = _makefile_var("ARCH")
    want = arch + flags
    # -Xptxas -v only produces the register report the Makefile archives
    want = [f for i, f in enumerate(want) if not (f == "-Xptxas" or (i and want[i - 1] == "-Xptxas"))]
    assert _rs_list("NVCCFLAGS") == want
    assert "-ffp-contract=off" in _rs_list("NVCCFLAGS")
    link = re.findall(r"cargo:rustc-link-lib=(?:\w+=)?(\w+)", _build_rs())
    mk = open(os.path.join(ROOT, "dawnsearch_b200", "csrc", "Makefile")).read()
    for lib in re.findall(r"-l(\w+)", mk.split("-shared", 1)[1]):
        assert lib in link, f"Makefile links -l{lib}, build.rs does not"
    # NCCL is bound at run time (dlopen in dawn_multi.cu: a DT_NEEDED on the system libnccl.so.2 clashes with the copy
    # torch bundles), so neither link line names it, and the library must not carry it as a dependency
    assert "nccl" not in link and "dl" in link
    assert "dlopen(\"libnccl.so.2\"" in open(os.path.join(ROOT, "dawnsearch_b200", "csrc", "dawn_multi.cu")).read()


def test_library_has_no_load_time_dependency_on_nccl(dawn):
    out = subprocess.run(["readelf", "-d", dawn.index.LIB_PATH], capture_output=True, text=True).stdout
    needed = re.findall(r"\(NEEDED\).*\[(.*?)\]", out)
    assert needed and not any("nccl" in n for n in needed), needed


def test_every_symbol_the_shim_binds_exists(dawn):
    shim = open(os.path.join(ROOT, "rust", "src", "index", "gpu_index.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', shim, flags=re.S).group(1)
    bound = set(re.findall(r"fn (dawn_[a-z0-9_]+)\(", block))
    assert len(bound) >= 30
    header = open(os.path.join(ROOT, "include", "dawn_index.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dawn_[a-z0-9_]+)\s*\(", header))
    assert bound <= declared, bound - declared
    out = subprocess.run(["nm", "-D", "--defined-only", dawn.index.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dawn_[a-z0-9_]+)", out))
    assert bound <= exported, bound - exported


def test_shim_mirrors_the_usearch_surface_the_reference_calls():
    shim = open(os.path.join(ROOT, "rust", "src", "index", "gpu_index.rs")).read()
    # search_provider.rs:102,115,117,133,149,178,214,246,280-284
    for m in ("pub fn new_index", "pub fn reserve", "pub fn add", "pub fn search", "pub fn size", "pub fn capacity",
              "pub fn save", "pub fn load", "pub fn view", "pub struct Matches", "pub struct IndexOptions"):
        assert m in shim, m
    for field in ("dimensions", "metric", "quantization", "connectivity", "expansion_add", "expansion_search"):
        assert re.search(rf"pub {field}:", shim), field
