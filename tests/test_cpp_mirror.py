"""include/gar.hpp (the compiled-language mirror of the Go API) builds against include/gar.h and behaves."""
import subprocess

from helpers import G, ROOT


def test_cpp_mirror_compiles_links_and_maps_errors(tmp_path):
    exe = tmp_path / "cpp_mirror_check"
    lib_dir = G.lib_path().parent
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp_mirror_check.cpp"),
                    "-L", str(lib_dir), "-lgar_b200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True)
    assert "cpp mirror ok" in out.stdout
