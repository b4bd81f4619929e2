"""The process boundary of the reference: ``ACE <param file>``.

The reference writes a parameter file (``pyaceqd/general_system/general_system.py:227-290``), runs
the external ``ACE`` binary on it (``:339-341``) and reads the text ``outfile`` back (``:342``).
This module implements that contract on the CUDA engine so that the UNMODIFIED reference can run
against it: put ``scripts/ACE`` on ``$PATH``.  Three pieces:

* :func:`write_param_file` -- the simulation parameter file, key for key in the reference's order
  (used by ``system_ace_stream(prepare_only=True)`` and by the round-trip tests);
* :func:`parse_param_file` / :func:`problem_from_params` -- the reader: ``key value...`` lines,
  operator expressions inside ``{ }`` (grammar of SURVEY App. B);
* :func:`main` -- the executable: propagation files are run on the engine and written as
  ``t Re Im Re Im ...`` rows with ``set_precision`` significant digits; process-tensor generation
  files (``write_PT``, ``:162-190``) build the Gaussian-bath PT with the host builder.
"""
from __future__ import annotations

import os
import re
import sys
from typing import Dict, List

import numpy as np

from pyaceqd_b200.jobs import FieldTable, Job
from pyaceqd_b200.problem import MTO, Problem, build_problem
from pyaceqd_b200.process_tensor import ProcessTensor

_BRACED = re.compile(r"\{([^{}]*)\}")


def write_param_file(path, *, dt, t_start, t_end, dict_zero="16", precision="12", pt_file=None, initial=None,
                     system_op=None, rf_op=None, rf_file=None, lindblad_ops=None, interaction_ops=None,
                     pulse_file_x=None, pulse_file_y=None, multitime_op=None, output_ops=(), out_file="ACE.out"):
    """Simulation parameter file with the keys and order of the reference (``:229-290``)."""
    lines = ["dt    {}".format(dt), "ta    {}".format(t_start), "te    {}".format(t_end),
             "dict_zero 1e-{}".format(dict_zero), "set_precision {}".format(precision),
             "use_symmetric_Trotter true"]
    if pt_file is not None:
        lines.append("add_PT    {}".format(pt_file))
    if initial is not None:
        lines.append("initial    {{ {} }}".format(initial))
    for op in system_op or []:
        lines.append("add_Hamiltonian {{ {} }}".format(op))
    if rf_op is not None:
        lines.append("add_Pulse file {} {{ -0.5*hbar*({}) }}".format(rf_file, rf_op))
    for op, rate in lindblad_ops or []:
        lines.append("add_Lindblad {}  {{ {} }}".format(rate, op))
    for op, pol in interaction_ops or []:
        lines.append("add_Pulse file {}  {{ -0.5*pi*hbar*({}) }}".format(pulse_file_y if pol == "y" else pulse_file_x, op))
    for m in multitime_op or []:
        lines.append("apply_Operator{applyFrom} {time} {{ {operator} }} {applyBefore}".format(**m))
    for op in output_ops:
        lines.append("add_Output {{ {} }}".format(op))
    lines.append("outfile {}".format(out_file))
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def parse_param_lines(lines) -> Dict[str, list]:
    """``{key: [entry, ...]}`` in file order; every entry is ``(words outside braces, [brace contents])``.
    The raw lines are kept under ``"__lines__"`` (operator order matters for coinciding times)."""
    params: Dict[str, list] = {"__lines__": []}
    for raw in lines:
        line = raw.split("#")[0].strip()
        if not line:
            continue
        params["__lines__"].append(line)
        braces = [b.strip() for b in _BRACED.findall(line)]
        words = _BRACED.sub(" ", line).split()
        params.setdefault(words[0], []).append((words[1:], braces))
    return params


def parse_param_file(path) -> Dict[str, list]:
    with open(path) as fh:
        return parse_param_lines(fh.readlines())


def _one(params, key, default=None):
    return params[key][0][0][0] if key in params else default


def _read_table(path) -> FieldTable:
    data = np.loadtxt(path, ndmin=2)
    t = data[:, 0]
    return FieldTable(float(t[0]), float(t[1] - t[0]) if len(t) > 1 else 1.0, data[:, 1] + 1j * data[:, 2])


def problem_from_params(params, coupling_diag=None):
    """``(Problem, tables, mto dicts)`` from parsed simulation parameters.  Pulse files map to drive
    tables in order of first appearance (at most three distinct files: x, y and rotating frame)."""
    slots, tables = {}, {}
    pulse_ops = []
    for words, braces in params.get("add_Pulse", []):
        if words[0] != "file" or not braces:
            raise ValueError("only 'add_Pulse file <path> { op }' is supported")
        f = words[1]
        if f not in slots:
            if len(slots) == 3:
                raise ValueError("more than three distinct pulse files")
            slots[f] = ("x", "y", "rf")[len(slots)]
            tables[slots[f]] = _read_table(f)
        pulse_ops.append((braces[0], slots[f]))
    dz = _one(params, "dict_zero", "1e-16")
    prob = build_problem(
        system_op=[b[0] for _, b in params.get("add_Hamiltonian", [])],
        initial=params["initial"][0][1][0] if "initial" in params else None,
        lindblad_ops=[(b[0], float(w[0])) for w, b in params.get("add_Lindblad", [])],
        raw_pulse_ops=pulse_ops, output_ops=[b[0] for _, b in params.get("add_Output", [])],
        dict_zero=float(dz), coupling_diag=coupling_diag)
    mtos = []
    for key, entries in params.items():
        if key == "__lines__" or not key.startswith("apply_Operator"):
            continue
        for words, braces in entries:
            mtos.append({"operator": braces[0], "time": float(words[0]), "applyFrom": key[len("apply_Operator"):],
                         "applyBefore": words[1] if len(words) > 1 else "false"})
    # file order matters for coinciding times (timebin/twophoton_new.py:436-438): re-read the order
    return prob, tables, mtos


def _mtos_in_file_order(params, prob: Problem) -> List[MTO]:
    out = []
    for line in params["__lines__"]:
        if not line.startswith("apply_Operator"):
            continue
        braces = _BRACED.findall(line)
        words = _BRACED.sub(" ", line).split()
        out.append({"operator": braces[0].strip(), "time": float(words[1]),
                    "applyFrom": words[0][len("apply_Operator"):],
                    "applyBefore": words[2] if len(words) > 2 else "false"})
    return prob.parse_mtos(out)


def setup_from_params(params):
    """``(Problem, ProcessTensor | None, Job)`` of a parsed propagation file."""
    pt = None
    coupling = None
    if "add_PT" in params:
        pt = ProcessTensor.load(_one(params, "add_PT"))
        coupling = (pt.meta or {}).get("coupling_diag")
        if coupling is None:
            raise ValueError("process tensor file carries no coupling operator; rebuild it with this engine")
    prob, tables, _ = problem_from_params(params, coupling_diag=coupling)
    job = Job(float(_one(params, "ta")), float(_one(params, "te")), float(_one(params, "dt")), tables=tables,
              mtos=_mtos_in_file_order(params, prob))
    return prob, pt, job


def write_outfile(params, job, out):
    """``t Re Im Re Im ...`` rows with ``set_precision`` significant digits (what ``:342`` parses)."""
    rows = np.empty((out.shape[1], 1 + 2 * out.shape[0]))
    rows[:, 0] = job.times()
    rows[:, 1::2] = out.real.T
    rows[:, 2::2] = out.imag.T
    np.savetxt(_one(params, "outfile", "ACE.out"), rows, fmt="%.{}g".format(int(_one(params, "set_precision", 12))),
               delimiter=" ")
    return rows


def _generate_pt(params, path):
    """Process-tensor generation file (``:162-190``): QDPhonon bath for the diagonal ``Boson_SysOp``."""
    from pyaceqd_b200.opparser import parse_operator
    from pyaceqd_b200.pt_builder import build_qd_phonon_pt
    from pyaceqd_b200.pt_builder import build_pt_from_spectral_density_file, write_spectral_density
    if "Boson_J_print" in params:          # <file> e_min e_max n
        w = params["Boson_J_print"][0][0]
        a_e_, a_h_ = float(_one(params, "Boson_J_a_e", 5.0)), _one(params, "Boson_J_a_h")
        write_spectral_density(w[0], a_e=a_e_, a_h=None if a_h_ is None else float(a_h_), e_min=float(w[1]) if len(w) > 1 else 0.0,
                               e_max=float(w[2]) if len(w) > 2 else 15.0, n=int(w[3]) if len(w) > 3 else 2000)
    if "Boson_J_from_file" in params:
        # the reference writes no Boson_SysOp in this mode (general_system.py:178-179): a two-level |1><1| coupling
        op = parse_operator(params["Boson_SysOp"][0][1][0]) if "Boson_SysOp" in params else np.diag([0.0, 1.0])
        dt = float(_one(params, "dt"))
        t_mem = float(_one(params, "t_mem", float(_one(params, "te", 2 * 20.48)) / 2))
        pt = build_pt_from_spectral_density_file(_one(params, "Boson_J_from_file"), np.real(np.diag(op)), dt, t_mem,
                                                 float(_one(params, "temperature", 4)), threshold=float(_one(params, "threshold", 1e-8)),
                                                 e_max=float(_one(params, "Boson_E_max", 7)))
        target = _one(params, "write_PT")
        pt.save(target)
        with open(target + "_initial", "w") as fh:
            fh.write("aceqd-b200 process tensor: see {}\n".format(os.path.basename(target)))
        return target
    op = parse_operator(params["Boson_SysOp"][0][1][0])
    a_e = float(_one(params, "Boson_J_a_e", 5.0))
    a_h = _one(params, "Boson_J_a_h")
    dt = float(_one(params, "dt"))
    infinite = str(_one(params, "use_Gaussian_infinite", "false")).lower() == "true"
    t_mem = float(_one(params, "t_mem", float(_one(params, "te", 2 * 20.48)) / 2))
    pt = build_qd_phonon_pt(coupling_diag=np.real(np.diag(op)), dt=dt, t_mem=t_mem, a_e=a_e,
                            a_h=None if a_h is None else float(a_h), temperature=float(_one(params, "temperature", 4)),
                            threshold=float(_one(params, "threshold", 1e-8)), e_max=float(_one(params, "Boson_E_max", 7)),
                            use_infinite=infinite)
    target = _one(params, "write_PT")
    pt.save(target)
    for suffix in ("_initial",):       # the reference tests for <pt_file>_initial (:156)
        with open(target + suffix, "w") as fh:
            fh.write("aceqd-b200 process tensor: see {}\n".format(os.path.basename(target)))
    return target


def run_param_file(path, engine=None):
    """Execute one parameter file; returns the output array ``[n_rows, 1 + 2 n_out]`` it wrote (or the
    PT path for a generation file)."""
    params = parse_param_file(path)
    if "write_PT" in params:
        return _generate_pt(params, path)
    prob, pt, job = setup_from_params(params)
    if engine is None:
        import pyaceqd_b200.engine as _engine
        engine = _engine.default_engine()
    return write_outfile(params, job, engine.run_jobs(prob, pt, [job])[0])


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        sys.stderr.write("usage: ACE <parameter file>\n")
        return 2
    try:
        run_param_file(argv[0])
    except Exception as exc:      # non-zero exit -> CalledProcessError in the reference (:339-341)
        sys.stderr.write("ACE (aceqd-b200): {}: {}\n".format(type(exc).__name__, exc))
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
