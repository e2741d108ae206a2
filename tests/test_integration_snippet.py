"""INTEGRATION.md's reference-side binding must compile against the reference's real data model: the snippet is
extracted from the document and compiled (-fsyntax-only) against a stub that declares `class individuals` and
`class Structure` with the reference's member names and types (fish_mod.cpp:23-25,35-62)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the reference's declarations (member names, types and array shapes of fish_mod.cpp:23-25,35-62; bodies omitted)
STUB = r"""
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>
#define num 60
#define n_cols 250
#define n_rows 250
class individuals {
public:
int n; int dim; int lag[num]; int colors[num][3]; int nfollow; int sample_rate[num]; int food_intake[num];
double speed[num]; double coords[num][2]; double dir[num]; bool sample[num]; double speed_factor[num];
};
class Structure {
public:
int latitude, longitude, resolution, res_factor, rows, cols, h, w;
};
"""


def snippet():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"## The binding a maintainer adds to fish_mod.cpp.*?```cpp\n(.*?)```", text, re.S)
    assert m, "INTEGRATION.md lost its binding snippet"
    return m.group(1)


def test_integration_snippet_compiles_against_the_reference_data_model(tmp_path):
    code = snippet()
    head, _, loop = code.partition("// in the step loop")
    assert loop, "snippet lost its step-loop part"
    body = "\n".join(l for l in loop.splitlines()[1:])
    src = (STUB + head + "\nint main() {\n  static individuals fish; static int quality[n_rows][n_cols]; Structure struc = Structure();\n"
           "  abm_attach(fish, quality, struc, num);\n" + body + "\n  return 0;\n}\n")
    f = tmp_path / "snippet.cpp"
    f.write_text(src)
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(f)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_host_model_keeps_speed_per_fish():
    """individuals::speed is an array (fish_mod.cpp:43), not a scalar"""
    hpp = open(os.path.join(ROOT, "fish_abm_b200", "host", "fishabm_host.hpp")).read()
    assert re.search(r"std::vector<double>\s+speed;", hpp)
    assert not re.search(r"double\s+speed\s*=", hpp)
