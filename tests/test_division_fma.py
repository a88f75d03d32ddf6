"""The clip path of the kernels divides through a correctly rounded reciprocal and two FMA corrections (``ddiv_rcp``,
csrc/gpr_device.cuh) instead of the division sequence.  The same operations on the CPU (C ``fma``) must reproduce IEEE division
bit for bit — the reference divides with NumPy (utils.py ensure_max_dyn_val, planning:610-645)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fma_corrected_division_is_correctly_rounded(tmp_path):
    exe = str(tmp_path / 'division_fma')
    # -ffp-contract=off: only the explicit fma() calls fuse; -mfma where the CPU has it (else libm's exact software fma)
    flags = ['-O2', '-ffp-contract=off']
    if 'fma' in open('/proc/cpuinfo').read().split():
        flags.append('-mfma')
    subprocess.run(['gcc', *flags, '-o', exe, os.path.join(HERE, 'division_fma.c'), '-lm'], check=True)
    out = subprocess.run([exe, '5000000'], capture_output=True, text=True, check=True, timeout=300)
    assert int(out.stdout.strip()) == 0, f'{out.stdout.strip()} quotients differ from IEEE division'
