"""The public structs of include/<code> have the reference's layout field for field: the same C program, once
compiled against our headers and once against the reference's, prints identical sizeof / offsetof tables
(objects compiled against either header set interoperate).  Needs the reference tree."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ZPIC_REFERENCE", "/root/reference")

FIELDS = {
    "em2d": {
        "t_part": "ix iy x y ux uy uz",
        "t_density": "n type start end custom_x custom_data_x custom_y custom_data_y custom_x_total_part custom_x_total_q",
        "t_species": "name part np np_max m_q energy q ppc density ufl uth nx dx box dt iter moving_window n_move n_sort",
        "t_smooth": "xtype ytype xlevel ylevel",
        "t_current": "J J_buf nx nrow gc box dx smooth dt iter moving_window",
        "t_emf_ext_fld": "E_type B_type E_0 B_0 E_custom B_custom E_custom_data B_custom_data E_part_buf B_part_buf",
        "t_emf_init_fld": "E_type B_type E_0 B_0 E_custom B_custom E_custom_data B_custom_data",
        "t_emf": "E B E_buf B_buf E_part B_part nx nrow gc box dx dt iter moving_window n_move ext_fld",
        "t_emf_laser": "type start fwhm rise flat fall a0 omega0 polarization W0 focus axis",
        "t_simulation": "dt tmax ndump n_species species emf current moving_window",
        "t_zdf_file": "fp mode ndatasets",
        "t_zdf_dataset": "name data_type ndims count data id offset",
        "t_zdf_chunk": "count start stride data",
        "t_zdf_grid_axis": "name type min max label units",
        "t_zdf_grid_info": "name ndims count label units axis",
        "t_zdf_iteration": "name n t time_units",
        "t_zdf_part_info": "name label np nquants quants qlabels qunits",
        "t_zdf_track_info": "name label ntracks ndump niter nquants quants qlabels qunits",
    },
    "em1d": {
        "t_part": "ix x ux uy uz",
        "t_species": "name part np np_max m_q energy q ppc density ufl uth nx dx box dt iter moving_window n_move n_sort bc_type",
        "t_smooth": "xtype xlevel",
        "t_current": "J J_buf nx gc box dx smooth dt iter bc_type",
        "t_emf": "E B E_buf B_buf E_part B_part nx gc box dx dt iter moving_window n_move bc_type ext_fld",
        "t_emf_laser": "start fwhm rise flat fall a0 omega0 polarization",
        "t_simulation": "dt tmax ndump n_species species emf current moving_window",
    },
}


def _table(code, include_dir, tmp_path, tag):
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "zpic.h"', '#include "simulation.h"', '#include "zdf.h"',
             "int main(void) {"]
    for t, fields in FIELDS[code].items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (t, t))
        for f in fields.split():
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (t, f, t, f))
    lines += ["return 0; }"]
    src = tmp_path / ("layout_%s.c" % tag)
    src.write_text("\n".join(lines))
    exe = tmp_path / ("layout_%s" % tag)
    r = subprocess.run(["gcc", "-std=gnu99", "-w", "-I" + include_dir, str(src), "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    return subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout


@pytest.mark.parametrize("code", ["em2d", "em1d"])
def test_struct_layouts_equal_the_reference(code, tmp_path):
    if not os.path.isdir(os.path.join(REF, code)):
        pytest.skip("reference tree not present")
    ours = _table(code, os.path.join(REPO, "include", code), tmp_path, "ours")
    ref = _table(code, os.path.join(REF, code), tmp_path, "ref")
    assert ours == ref
    assert len(ours.splitlines()) > 40
