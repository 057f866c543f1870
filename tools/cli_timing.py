"""End-to-end CLI wall time on the GPU box, BASELINE config 2/3 (44.1 kHz 16-bit stereo, 10 min):
the reference CLI (oracle/_ref/sela_ref_cli, all host cores) beside bin/sela and bin/sela_refmain
(the reference's main.cpp compiled unchanged over the GPU path).  Files live in /dev/shm so the disk
is not what is timed.  Writes gpurun_out/cli_timing_<tag>.txt; checks byte equality of every output."""
import filecmp, os, shutil, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sela_b200.synth import sine_noise
from sela_b200.wavio import write_wav

tag = sys.argv[1] if len(sys.argv) > 1 else "r"
D = "/dev/shm/sela_cli"
os.makedirs(D, exist_ok=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
lines = []
def say(s):
    print(s, flush=True)
    lines.append(s)
write_wav(f"{D}/in.wav", sine_noise(44100, 2, 600, seed=1), 44100)
env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "sela_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
def run(label, *cmd, extra=None):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=dict(env, **(extra or {})), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, cwd=ROOT)
    dt = time.perf_counter() - t0
    say("%-26s rc=%d  %7.3f s" % (label, r.returncode, dt))
    for line in r.stderr.decode().splitlines()[-12:]:
        say("      " + line)
REF, OURS, RM = "oracle/_ref/sela_ref_cli", "sela_b200/host/bin/sela", "sela_b200/host/bin/sela_refmain"
say("host cores: %d" % os.cpu_count())
say("GPUs visible: " + subprocess.run("nvidia-smi -L | wc -l; echo CUDA_VISIBLE_DEVICES=$CUDA_VISIBLE_DEVICES; nvidia-smi --query-gpu=persistence_mode --format=csv,noheader | head -1", shell=True, capture_output=True, text=True).stdout.replace("\n", " "))
run("reference -e", REF, "-e", f"{D}/in.wav", f"{D}/ref.sela")
run("reference -d", REF, "-d", f"{D}/ref.sela", f"{D}/ref.wav")
T = {"SELA_B200_TIMING": "1"}
for rep in (1, 2):
    run(f"sela (GPU) -e #{rep}", OURS, "-e", f"{D}/in.wav", f"{D}/ours.sela", extra=T)
    run(f"sela (GPU) -d #{rep}", OURS, "-d", f"{D}/ours.sela", f"{D}/ours.wav", extra=T)
run("sela classic (GPU) -e", OURS, "-e", f"{D}/in.wav", f"{D}/cl.sela", extra={"SELA_B200_CLASSIC": "1"})
run("sela classic (GPU) -d", OURS, "-d", f"{D}/cl.sela", f"{D}/cl.wav", extra={"SELA_B200_CLASSIC": "1"})
run("refmain (GPU) -e", RM, "-e", f"{D}/in.wav", f"{D}/rm.sela")
run("refmain (GPU) -d", RM, "-d", f"{D}/rm.sela", f"{D}/rm.wav")
for f in ("ours", "cl", "rm"):
    for ext in ("sela", "wav"):
        same = os.path.exists(f"{D}/{f}.{ext}") and filecmp.cmp(f"{D}/ref.{ext}", f"{D}/{f}.{ext}", shallow=False)
        say(f"{f}.{ext} == ref.{ext}: {same}")
for n in sorted(os.listdir(D)):
    say("%-10s %d bytes" % (n, os.path.getsize(f"{D}/{n}")))
shutil.rmtree(D)
open(os.path.join(ROOT, "gpurun_out", f"cli_timing_{tag}.txt"), "w").write("\n".join(lines) + "\n")
