"""CPU: the library's switch set built from the reference's own SHIPPED parameter.h equals the one built from the
condensed miluphcuda_b200/configs/<config>/parameter.h the tests use -- i.e. a maintainer who compiles
libb200sph against the parameter.h of examples/impact (etc.) gets the library the parity tests exercised.
The shipped files are staged by oracle/build_ref.sh under oracle/_ref/fixtures/ (they travel with the snapshot);
csrc/switches.h also rejects out-of-scope switches at compile time, so this compiles it with both files."""
import os
import subprocess

import pytest

import common

FIXTURES = os.path.join(common.REPO, "oracle", "_ref", "fixtures")
CSRC = os.path.join(common.REPO, "miluphcuda_b200", "csrc")

PROGRAM = r"""
#include <stdio.h>
#define B200SPH_NO_EOS_ENUM
#include "switches.h"
int main(void) {
#define X(name) printf("%s=%d;", #name, (int)(name));
    B200SPH_SWITCH_LIST(X)
#undef X
    return 0;
}
"""


def switch_string(include_dir, tmp_path, tag):
    src = tmp_path / f"sw_{tag}.c"
    exe = tmp_path / f"sw_{tag}"
    src.write_text(PROGRAM)
    subprocess.check_call(["gcc", "-I", include_dir, "-I", CSRC, str(src), "-o", str(exe)])
    return subprocess.check_output([str(exe)]).decode()


@pytest.mark.parametrize("config", common.CONFIGS)
def test_condensed_parameter_h_equals_shipped(config, tmp_path):
    shipped = os.path.join(FIXTURES, config)
    if not os.path.exists(os.path.join(shipped, "parameter.h")):
        pytest.skip("shipped parameter.h not staged (oracle/build_ref.sh needs /root/reference)")
    ours = switch_string(os.path.join(common.REPO, "miluphcuda_b200", "configs", config), tmp_path, "condensed")
    theirs = switch_string(shipped, tmp_path, "shipped")
    assert ours == theirs
