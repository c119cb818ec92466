import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from amplipy_b200 import alnio, cli, synth
g, prim, b = bench.make_workload(1_000_000, 2, "illumina")
d = tempfile.mkdtemp()
j = lambda n: os.path.join(d, n)
hdr = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:%d\n@PG\tID:synth\tPN:synth\n" % len(g)
alnio.write_bam(j("in.bam"), hdr, [("ref", len(g))], b)
synth.write_bed(j("primers.bed"), [(s, e, "p%d" % k) for k, (s, e) in enumerate(prim)])
synth.write_fasta(j("ref.fas"), "ref", g)
# warm run (engine / CUDA context set-up), then the profiled one
cli.main(["aio", "-i", j("in.bam"), "-p", j("primers.bed"), "-r", j("ref.fas"), "-ot", j("t0.bam"), "-ov", j("v0.vcf"), "-oc", j("c0.fas")])
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
cli.main(["aio", "-i", j("in.bam"), "-p", j("primers.bed"), "-r", j("ref.fas"), "-ot", j("t1.bam"), "-ov", j("v1.vcf"), "-oc", j("c1.fas")])
pr.disable()
print("second run seconds", time.perf_counter() - t0, file=sys.stderr)
pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(28)
