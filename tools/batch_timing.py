"""BASELINE config 5 shape, file to file: a batch of independent 3-minute 44.1 kHz stereo WAVs through
`sela -E` / `sela -D` (ONE process, files in /dev/shm), beside the reference CLI run once per file on a
few of them (all host cores each time).  Usage: python tools/batch_timing.py [n_files] [n_reference_files]
Writes gpurun_out/batch_timing.json."""
import filecmp, json, os, shutil, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sela_b200.synth import sine_noise
from sela_b200.wavio import write_wav

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 4
D = "/dev/shm/sela_batch"
shutil.rmtree(D, ignore_errors=True)
for sub in ("in", "enc", "dec", "ref"):
    os.makedirs(f"{D}/{sub}")
distinct = min(n_files, 8)
for i in range(distinct):                      # seeds = file index, as SURVEY.md 8d names them
    write_wav(f"{D}/in/f{i:04d}.wav", sine_noise(44100, 2, 180, seed=i), 44100)
for i in range(distinct, n_files):             # more files than seeds: byte copies (throughput only)
    shutil.copyfile(f"{D}/in/f{i % distinct:04d}.wav", f"{D}/in/f{i:04d}.wav")
wavs = [f"{D}/in/f{i:04d}.wav" for i in range(n_files)]
samples_per_file = (44100 * 180 // 2048) * 2048 * 2
env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "sela_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
SELA, REF = os.path.join(ROOT, "sela_b200/host/bin/sela"), os.path.join(ROOT, "oracle/_ref/sela_ref_cli")

def timed(*cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    return time.perf_counter() - t0, r

out = {"what": "config-5 shape: %d x 3-minute 44.1 kHz stereo WAV, file to file in /dev/shm" % n_files,
       "files": n_files, "samples": samples_per_file * n_files, "host_cores": os.cpu_count()}
for workers in (8, 16):
    e = dict(env, SELA_B200_WORKERS=str(workers))
    t0 = time.perf_counter()
    r = subprocess.run([SELA, "-E", f"{D}/enc"] + wavs, env=e, capture_output=True, text=True)
    te = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-500:]
    t0 = time.perf_counter()
    r = subprocess.run([SELA, "-D", f"{D}/dec"] + [f"{D}/enc/f{i:04d}.sela" for i in range(n_files)], env=e,
                       capture_output=True, text=True)
    td = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-500:]
    out["batch_workers_%d" % workers] = {
        "encode_s": round(te, 3), "decode_s": round(td, 3),
        "encode_MSamples_s": round(out["samples"] / te / 1e6, 1), "decode_MSamples_s": round(out["samples"] / td / 1e6, 1)}
# steady-state phases of one worker thread (SELA_B200_TIMING), last file of six
e = dict(env, SELA_B200_WORKERS="1", SELA_B200_TIMING="1")
r = subprocess.run([SELA, "-E", f"{D}/enc"] + wavs[:6], env=e, capture_output=True, text=True)
out["one_worker_encode_phases_last_file"] = [l.replace("[sela_b200] ", "").split() for l in r.stderr.splitlines()[-4:]]
r = subprocess.run([SELA, "-D", f"{D}/dec"] + [f"{D}/enc/f{i:04d}.sela" for i in range(6)], env=e, capture_output=True, text=True)
out["one_worker_decode_phases_last_file"] = [l.replace("[sela_b200] ", "").split() for l in r.stderr.splitlines()[-5:]]
# reference CLI, one process per file
te = td = 0.0
same = True
for i in range(n_ref):
    dt, r = timed(REF, "-e", wavs[i], f"{D}/ref/f{i:04d}.sela"); te += dt
    dt, r = timed(REF, "-d", f"{D}/ref/f{i:04d}.sela", f"{D}/ref/f{i:04d}.wav"); td += dt
    same &= filecmp.cmp(f"{D}/ref/f{i:04d}.sela", f"{D}/enc/f{i:04d}.sela", shallow=False)
    same &= filecmp.cmp(f"{D}/ref/f{i:04d}.wav", f"{D}/dec/f{i:04d}.wav", shallow=False)
out["reference_cli"] = {"files": n_ref, "encode_s_per_file": round(te / n_ref, 3), "decode_s_per_file": round(td / n_ref, 3),
                        "encode_MSamples_s": round(samples_per_file * n_ref / te / 1e6, 1),
                        "decode_MSamples_s": round(samples_per_file * n_ref / td / 1e6, 1)}
out["outputs_identical_to_reference_cli"] = bool(same)
# every decoded file equals its source (whole frames; these lengths are not frame multiples -> compare the prefix)
ok = True
for i in range(min(n_files, 8)):
    a, b = open(wavs[i], "rb").read(), open(f"{D}/dec/f{i:04d}.wav", "rb").read()
    ok &= a[44:44 + len(b) - 44] == b[44:]
out["decoded_equals_source_prefix"] = bool(ok)
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "batch_timing.json"), "w").write(json.dumps(out, indent=1) + "\n")
shutil.rmtree(D)
