import sys, os, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_renderer_b200 import renderer as R, scenes
from oracle import oracle
def run(sc, compare=True):
    rnd = R.Renderer(); rnd.resize_internal_buffers(sc.width, sc.height)
    batch = R.ShapeBatch(rnd, sc.dynamic_stroke_options, sc.paths, sc.shape_path_begin)
    cmds = scenes.stencil_cover_commands(sc.n_shapes)
    rp = rnd.begin_render_pass(); rp.set_instances(sc.transforms(), sc.colors); rp.render_batch(batch, cmds); rp.submit()
    color, stencil = rnd.read_color(), rnd.read_stencil()
    st = rnd.stats()
    res = np.argwhere(stencil[..., 0] != 0)
    print(sc.name, sc.paths.n_paths, "prims", st.primitives, "pairs", st.tile_pairs, "covered", st.covered_samples, "residue", len(res), res[:8].tolist())
    if compare:
        t = time.time()
        refs = [oracle.shape_from_paths(sc.dynamic_stroke_options, sc.paths, int(sc.shape_path_begin[i]), int(sc.shape_path_begin[i + 1])) for i in range(sc.n_shapes)]
        bad_shapes = [i for i in range(sc.n_shapes) if not (np.array_equal(batch[i].vertex_buffer(), refs[i].vertex_buffer) and np.array_equal(batch[i].index_buffer(), refs[i].index_buffer))]
        print("  tess mismatching shapes:", bad_shapes[:10], "oracle tess s", time.time() - t)
        ocmds = [(int(c[0]), int(c[1]), int(c[2]), int(c[3]), 0, 0, 0) for c in cmds]
        rc, rs, _, cov = oracle.render(rnd.config.to_c(), sc.width, sc.height, refs, ocmds, sc.transforms(), sc.colors, threads=oracle.max_threads())
        d = np.argwhere(rs != stencil)
        dc = np.argwhere((rc.view(np.uint32) != color.view(np.uint32)).any(-1))
        print("  oracle covered", cov, "stencil diffs", len(d), d[:6].tolist(), "colour diffs", len(dc), dc[:6].tolist(), "oracle residue", int((rs != 0).sum()))
        if len(dc):
            ys, xs = dc[:, 0] // 16, dc[:, 1] // 16
            tiles = np.unique(ys * 1000 + xs)
            print("  differing tiles:", len(tiles), tiles[:10].tolist())
    batch.close(); rnd.close()
run(scenes.glyph_like_fills(4000, extent=(1920, 1080), glyphs_per_shape=400))
run(scenes.glyph_like_fills(20000, extent=(1920, 1080), glyphs_per_shape=400))
run(scenes.glyph_like_fills(100000), compare=len(sys.argv) > 1)
