"""Control-file front end and case driver (fvens_b200/host/controlparser.hpp, casesolvers.hpp, fvens_steady.cpp)
against the reference's own control files (tests/golden/ctrl, copied from testcases/naca0012, tests/flow-general,
tests/inv-2dcyl): what parse_flow_controlfile (reference src/utilities/controlparser.cpp:60-290) must extract,
including Boost.PropertyTree INFO corner cases the reference's files rely on (#include, quoted and unquoted values,
`key` and `{` on different lines, a second word after a value starting a new key).
The end-to-end run of the fvens_steady executable on the GPU is in tests/test_gpu_steady_case.py."""
import json
import math
import os
import re
import subprocess
import xml.dom.minidom

import numpy as np
import pytest

from common import ROOT, MESHDIR

CTRL = os.path.join(ROOT, "tests", "golden", "ctrl")
PROBE = os.path.join(ROOT, "tests", "cpp", "test_controlparser")
STEADY = os.path.join(ROOT, "tests", "cpp", "fvens_steady")
BC = dict(slipwall=0, farfield=1, inflowoutflow=2, subsonic_inflow=3, extrapolation=4, periodic=5, isothermalwall=6, adiabaticwall=7)


def parse(name, *cmd):
    r = subprocess.run([PROBE, os.path.join(CTRL, name), *cmd], capture_output=True, text=True, timeout=60)
    return json.loads(r.stdout.strip().splitlines()[-1])       # the parser announces command-line overrides on stdout, as the reference


def test_naca0012_explicit_control_file():
    o = parse("naca0012-transonic-explicit.ctrl")
    assert o["meshfile"] == "testcases/naca0012/grids/NACA0012_inv.su2" and o["vtu"] == "naca.vtu" and o["logfile"] == "naca-log"
    assert o["lognres"] == 0 and o["flowtype"] == "EULER" and o["viscsim"] == 0 and o["sim_type"] == "STEADY"
    assert o["gamma"] == 1.4 and o["Minf"] == 0.8 and o["alpha"] == math.pi/180.0*1.25
    assert o["Tinf"] == 298.0 and o["Reinf"] is None and o["Pr"] is None          # the reference's 1/0 and 0/0 for Euler runs
    assert (o["invflux"], o["gradient"], o["limiter"], o["order2"]) == ("HLLC", "LEASTSQUARES", "WENO", 1)
    assert o["limiter_param"] == 20.0                                               # read here, lost in the reference
    assert o["pseudotimetype"] == "EXPLICIT" and (o["initcfl"], o["endcfl"], o["tolerance"], o["maxiter"]) == (0.2, 0.2, 1e-5, 200000)
    assert o["usestarter"] == 1 and (o["firstinitcfl"], o["firsttolerance"], o["firstmaxiter"]) == (0.8, 0.1, 10000)
    assert o["lwalls"] == [2] and o["surfnameprefix"] == "inv-naca" and o["vol_output_reqd"] == "NO"
    assert o["bcs"] == [dict(tag=2, type=BC["slipwall"], vals=[]), dict(tag=4, type=BC["farfield"], vals=[])]
    # firstorder_spatial_numerics_config / extract_spatial_numerics_config
    assert o["first_order"] == dict(gradient="NONE", limiter="NONE", order2=0, flux="HLLC") and o["main"] == dict(gradient="LEASTSQUARES", order2=1)
    assert o["phys_nbc"] == 2


def test_viscous_control_file_and_info_corner_cases():
    o = parse("flow-general-test.ctrl")
    assert o["flowtype"] == "NAVIERSTOKES" and o["viscsim"] == 1 and o["useconstvisc"] == 0
    assert (o["Minf"], o["Tinf"], o["Reinf"], o["Pr"]) == (0.5, 288.15, 5000.0, 0.72)
    assert (o["invflux"], o["gradient"], o["limiter"]) == ("ROE", "LEASTSQUARES", "NONE")
    bcs = {b["tag"]: b for b in o["bcs"]}
    assert bcs[4]["type"] == BC["farfield"] and bcs[2]["type"] == BC["adiabaticwall"] and bcs[3]["type"] == BC["isothermalwall"]
    assert bcs[2]["vals"] == [0.0]                      # quoted "0.0"
    # `boundary_values 0.0  290.0` is unquoted: Boost's INFO parser takes "0.0" as the value and starts a new key with
    # "290.0" - so does this parser (the reference's Isothermalwall2D then reads past the end of bc_vals)
    assert bcs[3]["vals"] == [0.0]


def test_include_directive_and_command_line_overrides():
    o = parse("expl-cyl-ls-hllc.ctrl", "source_dir", CTRL, "mesh_file", "some/mesh.msh", "log_file_prefix", "/tmp/x")
    assert o["meshfile"] == "some/mesh.msh" and o["logfile"] == "/tmp/x" and o["lognres"] == 1
    assert o["Minf"] == 0.38 and o["lwalls"] == [2] and o["surfnameprefix"] == "2dcyl"
    assert (o["invflux"], o["gradient"], o["limiter"], o["limiter_param"]) == ("HLLC", "LEASTSQUARES", "NONE", 1.0)
    assert (o["initcfl"], o["tolerance"], o["maxiter"], o["firstmaxiter"]) == (0.25, 1e-4, 10000, 1000)
    # without the source directory the include cannot be resolved: an error, not a silent default
    assert "cannot open control file" in parse("expl-cyl-ls-hllc.ctrl", "source_dir", "/nonexistent")["error"]


def test_implicit_options_are_parsed_but_missing_keys_are_errors(tmp_path):
    o = parse("naca0012-transonic-implicit.ctrl")
    assert o["pseudotimetype"] == "IMPLICIT" and o["invfluxjac"] == o["invflux"]       # "consistent"
    bad = tmp_path / "bad.ctrl"
    bad.write_text(open(os.path.join(CTRL, "naca0012-transonic-explicit.ctrl")).read().replace("freestream_Mach_number  0.8", ""))
    assert "No such node (flow_conditions.freestream_Mach_number)" in parse(str(bad))["error"]
    bad.write_text("io {\n mesh_file \"x\"\n")
    assert "unmatched" in parse(str(bad))["error"]


def test_fvens_steady_rejects_what_it_cannot_run(tmp_path):
    """No GPU needed: the driver validates the case before it touches the mesh or the device."""
    r = subprocess.run([STEADY, os.path.join(CTRL, "naca0012-transonic-implicit.ctrl")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "pseudotime_stepping_type implicit" in r.stderr
    r = subprocess.run([STEADY, str(tmp_path / "missing.ctrl")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "cannot open control file" in r.stderr
    r = subprocess.run([STEADY], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "control file" in r.stderr
    r = subprocess.run([STEADY, os.path.join(CTRL, "naca0012-transonic-explicit.ctrl"), "--bogus"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2


def test_reference_parse_test_against_its_testdata():
    """tests/utils/testparse.cpp (Utils_ParseControlInfo): inv-explicit.ctrl parsed and compared, field by field and
    in the order of parse_solution_file, with inv-explicit.testdata (both files are the reference's fixtures)."""
    o = parse("inv-explicit.ctrl")
    t = open(os.path.join(CTRL, "inv-explicit.testdata")).read().split()
    it = iter(t)
    assert o["meshfile"] == next(it) and o["vtu"] == next(it) and o["logfile"] == next(it) and o["lognres"] == int(next(it))
    assert o["flowtype"] == next(it) and o["gamma"] == float(next(it))
    assert o["alpha"] == float(next(it))*math.pi/180.0 and o["Minf"] == float(next(it))
    assert o["bcs"][0] == dict(tag=int(next(it)), type=BC["slipwall"], vals=[])
    assert o["bcs"][1] == dict(tag=int(next(it)), type=BC["farfield"], vals=[])
    assert len(o["lwalls"]) == int(next(it)) and o["surfnameprefix"] == next(it) and o["vol_output_reqd"] == next(it)
    assert o["sim_type"] == next(it) and o["invflux"] == next(it) and o["gradient"] == next(it) and o["limiter"] == next(it)
    assert o["pseudotimetype"] == next(it)
    assert (o["initcfl"], o["endcfl"], o["tolerance"], o["maxiter"]) == (float(next(it)), float(next(it)), float(next(it)), int(next(it)))
    assert (o["firstinitcfl"], o["firstendcfl"], o["firsttolerance"], o["firstmaxiter"]) == (float(next(it)), float(next(it)), float(next(it)), int(next(it)))
    assert next(it, None) is None
