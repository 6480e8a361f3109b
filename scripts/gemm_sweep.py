import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == 'x':
    from gpsurvival_b200 import gp
    fit = gp.Fit(0)
    for (M, N, K) in [(1000,1000,1000),(2048,2048,256),(2048,2048,2048),(8192,8192,8192)]:
        ms = fit.bench_gemm(M, N, K, 5)
        print(f"  tile={os.environ.get('GPS_TILE')} {M}x{N}x{K}: {ms*1e3:9.1f} us  {2.0*M*N*K/ms/1e9:6.2f} TF/s", flush=True)
else:
    for t in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("0", "3", "4")):
        env = dict(os.environ, GPS_TILE=t)
        subprocess.run([sys.executable, __file__, "x"], env=env)
