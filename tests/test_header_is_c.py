"""include/fishabm.h must be a plain C header (the drop-in boundary is a C ABI): compile a C99 translation unit that
uses every type and a few calls, and link-check it against the library."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "fishabm.h"
#include <stddef.h>
int use(void) {
    fishabm_params p;
    fishabm_ctx* ctx = NULL;
    fishabm_pass_stats st;
    fishabm_slab_info info;
    fishabm_slab_endpoint ep;
    int rc = fishabm_params_default(&p);
    (void)st; (void)info; (void)ep;
    if (rc != FISHABM_OK) return rc;
    p.n = 60; p.mode = FISHABM_MODE_CELL_FP32;
    rc = fishabm_create(&p, &ctx);
    if (rc == FISHABM_OK) { rc = fishabm_step(ctx, 1); fishabm_destroy(ctx); }
    return rc + (int)sizeof(fishabm_slab_endpoint) + FISHABM_UNIQUE_ID_BYTES + FISHABM_STREAM_SAMPLE;
}
int main(void) { return use() == 12345; }
'''


def test_header_compiles_as_c99_and_links(tmp_path):
    c = tmp_path / "use.c"
    c.write_text(SRC)
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-c", str(c), "-o", str(tmp_path / "use.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = os.path.join(ROOT, "fish_abm_b200")
    r = subprocess.run(["gcc", str(tmp_path / "use.o"), "-o", str(tmp_path / "use"), "-L", lib, "-lfishabm", f"-Wl,-rpath,{lib}",
                        "-Wl,-rpath-link,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
