import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_renderer_b200 import renderer as R, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ext = (3840, 2160) if len(sys.argv) > 2 and sys.argv[2] == "4k" else (1920, 1080)
sc = scenes.tiger_like(n, extent=ext, instance_px=(60.0, 260.0))
msaa = 4 if "msaa4" in sys.argv else 1
rnd = R.Renderer(R.Configuration(alpha_layer_count=2, msaa_sample_count=msaa)); rnd.resize_internal_buffers(sc.width, sc.height); rnd.enable_timing(True)
batch = R.ShapeBatch(rnd, [], sc.paths, sc.shape_path_begin)
for it in range(3):
    import time
    rp = rnd.begin_render_pass(); rp.set_instances(sc.transforms, sc.colors); sc.record(rp, batch)
    rnd.synchronize(); t0 = time.perf_counter(); rp.submit(); t1 = time.perf_counter(); rnd.synchronize(); t2 = time.perf_counter()
    st = rnd.stats()
    print(f"tiger_like({n}) {ext}: submit call {1e3*(t1-t0):.3f} ms, until done {1e3*(t2-t0):.3f} ms (host clock); bin {st.last_bin_ms:.3f} ms raster {st.last_raster_ms:.3f} ms pairs {st.tile_pairs} prims {st.primitives} covered {st.covered_samples}")
