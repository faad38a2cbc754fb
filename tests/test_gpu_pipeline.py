"""End-to-end checks of the host mirrors: Facet.score_images / BatchProcessor / ScoringPipeline vs the oracle."""
import numpy as np
import pytest

from facet_b200.synth import synth_embeddings, synth_image_bgr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scorer():
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.processing.scorer import Facet
    tags = synth_embeddings(24, seed=7, cluster_fraction=0.0)
    names = [f"tag{i // 2}" for i in range(24)]
    return Facet(random_state_dict(0), text_embeddings=tags, tag_names=names), tags, names


def test_batch_processor_mixed_shapes_and_errors(scorer):
    import torch
    from facet_b200.processing.batch_processor import BatchProcessor
    from oracle import cpu_port, technical_np as onp, vit_torch
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.models.tagger import select_tags
    sc, tags, names = scorer
    shapes = [(256, 384), (200, 320), (256, 384), (97, 131)]
    items = [{"path": f"/x/img{i}.jpg", "img_cv": synth_image_bgr(i, h, w)} for i, (h, w) in enumerate(shapes)]
    items.insert(2, {"path": "/x/broken.jpg", "error": "Failed to load image"})
    items.append({"path": "/x/none.jpg", "img_cv": None})
    bp = BatchProcessor(sc, batch_size=4)
    res = list(bp.process_items(items))
    assert [r.get("path") for r in res] == [it["path"] for it in items]
    assert "error" in res[2] and "error" in res[-1]
    sd = {k: v.cuda() for k, v in random_state_dict(0).items()}
    for it, r in zip(items, res):
        if "error" in r:
            continue
        img = it["img_cv"]
        m = onp.all_metrics(img, mono_threshold=0.10)
        assert r["histogram_data"] == m["histogram"]["histogram_bytes"]
        assert r["is_monochrome"] == m["monochrome"]["is_monochrome"]
        assert r["shadow_clipped"] == m["histogram"]["shadow_clipped"] and r["highlight_clipped"] == m["histogram"]["highlight_clipped"]
        assert abs(r["raw_sharpness_variance"] - m["sharpness"]["raw_variance"]) <= 1e-9 * max(1.0, m["sharpness"]["raw_variance"])
        assert r["noise_sigma"] == m["noise"]["noise_sigma"] and r["contrast_score"] == m["contrast"]["contrast_score"]
        assert r["dynamic_range_stops"] == m["dynamic_range"]["dynamic_range_stops"]
        from oracle import phash as oph
        assert r["phash"] == oph.phash_hex(img)
        clip_in = cpu_port.clip_preprocess_pil(img).unsqueeze(0).cuda()
        ref = vit_torch.score_batch(sd, clip_in, torch.from_numpy(tags).cuda())
        emb = np.frombuffer(r["clip_embedding"], np.float32)
        assert len(r["clip_embedding"]) == 3072
        assert float(np.dot(emb, ref["embedding"][0].cpu().numpy())) >= 0.999
        assert abs(r["aesthetic"] - round(float(ref["aesthetic"][0]), 2)) <= 0.011
        want_tags = select_tags(names, ref["tag_sims"][0].cpu().numpy(), 0.22, 5)
        got_tags = r["tags"].split(",") if r["tags"] else []
        sims = dict(zip(names, ref["tag_sims"][0].cpu().numpy()))
        # identical unless a similarity sits within rounding distance of the threshold
        if all(abs(float(v) - 0.22) > 2e-3 for v in sims.values()):
            assert set(got_tags) == set(want_tags)
    assert bp.metrics["images_failed"] == 2 and bp.metrics["images_processed"] == 4


def test_host_pipeline_matches_direct_call(scorer):
    import torch
    from facet_b200.processing.pipeline import ScoringPipeline
    sc, _, _ = scorer
    frames = np.stack([synth_image_bgr(10 + i, 256, 384) for i in range(11)])
    host = torch.from_numpy(frames).pin_memory()
    pipe = ScoringPipeline(sc, chunk=4)
    res = pipe.run_host(host)
    direct = sc.score_images_device(torch.from_numpy(frames).cuda())
    assert np.array_equal(res["hist256"], direct["hist256"].cpu().numpy().view(np.uint32))
    assert np.array_equal(res["sums"], direct["sums"].cpu().numpy())
    assert np.array_equal(res["phash"], direct["phash"].cpu().numpy().view(np.uint64))
    np.testing.assert_allclose(res["embedding"], direct["embedding"].cpu().numpy(), rtol=0, atol=1e-6)
    assert pipe.h2d_bytes(11, 256, 384) == frames.nbytes
    # consecutive host batches as one stream: same results as one call on their concatenation
    parts = [host[:5], host[5:6], host[6:]]
    streamed = pipe.run_host_stream(iter(parts))
    for key in ("hist256", "sums", "phash"):
        assert np.array_equal(streamed[key], res[key]), key
    np.testing.assert_allclose(streamed["embedding"], res["embedding"], rtol=0, atol=1e-6)
    # frames uploaded as decoded (stored rotated, EXIF code 6, RGB): the device orients them
    stored = np.ascontiguousarray(np.rot90(frames[..., ::-1], k=1, axes=(1, 2)))
    res6 = pipe.run_host(torch.from_numpy(stored).pin_memory(), rgb_order=True, orientation=6)
    assert np.array_equal(res6["hist256"], res["hist256"]) and np.array_equal(res6["sums"], res["sums"])
    assert np.array_equal(res6["phash"], res["phash"])
    np.testing.assert_allclose(res6["embedding"], res["embedding"], rtol=0, atol=1e-6)


def test_tagger_and_single_image_twins(scorer):
    from PIL import Image
    sc, tags, names = scorer
    img = synth_image_bgr(3, 300, 400)
    pil = Image.fromarray(np.ascontiguousarray(img[..., ::-1]))
    a, e, q, model = sc.get_aesthetic_and_quality(pil)
    (a2, e2, _, _), = sc.get_aesthetic_and_quality_batch([pil])
    assert model == "clip-mlp" and q is None and len(e) == 3072 and a == a2 and e == e2
    got = sc.tagger.get_tags_from_embedding(e, threshold=-1.0, max_tags=3)
    sims = np.frombuffer(e, np.float32) @ tags.T
    best = {}
    for n, s in zip(names, sims):
        best[n] = max(best.get(n, -9), float(s))
    want = [t for t, _ in sorted(best.items(), key=lambda kv: -kv[1])[:3]]
    assert got == want
    assert 0.0 <= sc.score_from_embedding(e) <= 10.0
    # tagger.py:116-158: scores of the tags above a threshold, and the art test on the ten best tags
    scores = sc.tagger.get_tags_with_scores(e, threshold=-1.0)
    assert set(scores) == set(best) and all(abs(scores[t] - round(best[t], 3)) <= 2e-3 for t in best)
    ten = set(sc.tagger.get_tags_from_embedding(e, threshold=-1.0, max_tags=10))
    sc.tagger.art_tags = {want[0]}
    assert sc.tagger.is_artwork(e, threshold=-1.0) is True
    sc.tagger.art_tags = {"no such tag"}
    assert sc.tagger.is_artwork(e, threshold=-1.0) is False and len(ten) == 10
    sc.tagger.art_tags = set()


def test_batch_processor_with_config_fills_aggregate_and_category(tmp_path):
    """With a ScoringConfig the pass also yields `category` / `aggregate` (processing/aggregate.py, pinned on the
    CPU against the reference); here: they are computed from the unrounded analyzer values of the device pass."""
    import json
    import os
    from facet_b200.config import ScoringConfig
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.processing.aggregate import calculate_aggregate_logic
    from facet_b200.processing.batch_processor import BatchProcessor
    from facet_b200.processing.scorer import Facet
    with open(os.path.join(os.path.dirname(__file__), "golden", "aggregate_golden.json")) as f:
        cfg_dict = json.load(f)["cases"][0]["config"]
    path = tmp_path / "scoring_config.json"
    path.write_text(json.dumps(cfg_dict))
    cfg = ScoringConfig(str(path))
    sc = Facet(random_state_dict(0), config=cfg)
    items = [{"path": f"/x/{i}.jpg", "img_cv": synth_image_bgr(40 + i, 128, 192)} for i in range(5)]
    items[1]["face_res"] = {"face_count": 1, "face_quality": 8.0, "eye_sharpness": 7.0, "is_blink": 0, "face_area": 4000,
                            "bbox": [60, 30, 120, 90], "face_sharpness": 900.0, "raw_eye_sharpness": 50.0,
                            "is_group_portrait": 0, "max_face_confidence": 0.95, "face_details": []}
    items[2]["exif_data"] = {"iso": 64, "f_stop": 1.8}
    res = list(BatchProcessor(sc, batch_size=8).process_items(items))
    direct = sc.score_images(np.stack([it["img_cv"] for it in items]), mono_threshold=cfg.get_monochrome_settings()["saturation_threshold_percent"] / 100)
    for it, r, d in zip(items, res, direct):
        assert "error" not in r, r
        assert r["config_version"] == cfg.version_hash and r["tags"] is None
        face = it.get("face_res")
        m = {"aesthetic": d["aesthetic_unrounded"], "face_count": face["face_count"] if face else 0,
             "face_quality": face["face_quality"] if face else 0, "eye_sharpness": face["eye_sharpness"] if face else 0,
             "tech_sharpness": d["tech_sharpness_unrounded"], "color_score": d["color_score_unrounded"],
             "exposure_score": d["exposure_score_unrounded"], "face_ratio": (4000 / (128 * 192)) if face else 0.0,
             "comp_score": r["comp_score"], "isolation_bonus": max(1.0, 900.0 / (d["raw_sharpness_variance"] + 1)) if face else 1.0,
             "is_blink": 0, "shadow_clipped": d["shadow_clipped"], "highlight_clipped": d["highlight_clipped"],
             "is_silhouette": 1 if (d["is_silhouette"] and face) else 0, "histogram_spread": d["histogram_spread"],
             "iso": it.get("exif_data", {}).get("iso"), "f_stop": it.get("exif_data", {}).get("f_stop"),
             "quality_score": None, "scoring_model": "clip-mlp"}
        want, cat = calculate_aggregate_logic(m, cfg)
        assert r["category"] == cat and r["aggregate"] == round(want, 2)
        assert 0.0 <= r["aggregate"] <= 10.0
    assert res[1]["category"] in ("portrait", "portrait_bw", "silhouette", "human_others")
    assert res[0]["category"] in ("default", "monochrome", "night")


def test_batch_processor_streamed_path_equals_the_blocking_one(scorer):
    """`process_items_streamed` (pinned staging, copy / compute / D2H streams, one packed record per image, host
    worker thread) returns exactly the dicts of `process_items`: mixed shapes, error items, chunks and ViT batches
    that do not divide the item count, pinned and pageable frames."""
    import torch
    from facet_b200.processing.batch_processor import BatchProcessor
    sc, tags, names = scorer
    shapes = [(256, 384)] * 5 + [(200, 320)] * 3 + [(256, 384)] * 2 + [(97, 131)]
    pinned = torch.empty((len(shapes), 256, 384, 3), dtype=torch.uint8, pin_memory=True)
    items = []
    for i, (h, w) in enumerate(shapes):
        img = synth_image_bgr(60 + i, h, w)
        if (h, w) == (256, 384) and i % 2 == 0:          # every other frame lives in pinned host memory
            view = pinned[i].numpy()
            view[:] = img
            img = view
        items.append({"path": f"/x/s{i}.jpg", "img_cv": img})
    items.insert(4, {"path": "/x/broken.jpg", "error": "Failed to load image"})
    items.append({"path": "/x/none.jpg", "img_cv": None})
    items.append("not a dict")
    want = list(BatchProcessor(sc, batch_size=16).process_items(items))
    bp = BatchProcessor(sc, batch_size=16)
    for chunk, vb in ((4, 6), (16, 64), (1, 1)):
        got = bp.process_items_streamed(items, chunk=chunk, vit_batch=vb)
        assert len(got) == len(want)
        for g, w_ in zip(got, want):
            assert g.keys() == w_.keys(), (g.keys() ^ w_.keys())
            for k in w_:
                assert g[k] == w_[k], (chunk, vb, w_.get("path"), k)
    assert bp.metrics["h2d_bytes"] == 3 * sum(h * w * 3 for h, w in shapes) and bp.metrics["d2h_bytes"] > 0


def test_streamed_path_takes_jpeg_file_bytes(scorer):
    """Items that carry the FILE BYTES (`jpeg`) are decoded on the device; results equal those of the same frames decoded
    by Pillow on the host (the reference's loader, utils/image_loading.py:90-106) and passed as `img_cv`, EXIF orientation
    included.  Progressive files use the item's `img_cv` or the host loader, corrupt ones become error items."""
    import io
    import torch
    from PIL import Image, ImageOps
    from facet_b200.processing.batch_processor import BatchProcessor
    sc, tags, names = scorer
    items_jpeg, items_raw = [], []
    for i, (h, w, kw, code) in enumerate([(256, 384, {"quality": 90, "restart_marker_blocks": 6}, 1), (256, 384, {"quality": 90, "restart_marker_blocks": 6}, 1),
                                          (200, 320, {"quality": 75, "subsampling": 0, "restart_marker_rows": 1}, 1),
                                          (256, 384, {"quality": 90, "restart_marker_blocks": 6}, 6), (120, 131, {"quality": 85}, 1)]):
        rgb = synth_image_bgr(70 + i, h, w)[:, :, ::-1].copy()
        ex = Image.Exif()
        ex[0x0112] = code
        buf = io.BytesIO()
        Image.fromarray(rgb).save(buf, "JPEG", exif=ex, **kw)
        data = buf.getvalue()
        decoded = np.asarray(ImageOps.exif_transpose(Image.open(io.BytesIO(data))).convert("RGB"))[:, :, ::-1].copy()
        pinned = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = np.frombuffer(data, np.uint8)
        items_jpeg.append({"path": f"/x/j{i}.jpg", "jpeg": pinned.numpy() if i % 2 else data, "_keep": pinned})
        items_raw.append({"path": f"/x/j{i}.jpg", "img_cv": decoded})
    # progressive stream + host-decoded frame, corrupt stream, progressive stream without a frame
    rgb = synth_image_bgr(90, 64, 96)[:, :, ::-1].copy()
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, "JPEG", progressive=True)
    items_jpeg.append({"path": "/x/prog.jpg", "jpeg": buf.getvalue(), "img_cv": rgb[:, :, ::-1].copy()})
    items_raw.append({"path": "/x/prog.jpg", "img_cv": rgb[:, :, ::-1].copy()})
    good = items_jpeg[0]["jpeg"]
    items_jpeg.append({"path": "/x/corrupt.jpg", "jpeg": good[:len(good) // 2] + b"\xff\xd9"})
    items_jpeg.append({"path": "/x/prog_only.jpg", "jpeg": buf.getvalue()})
    bp = BatchProcessor(sc, batch_size=16)
    want = bp.process_items_streamed(items_raw, chunk=4, vit_batch=8)
    got = bp.process_items_streamed(items_jpeg, chunk=4, vit_batch=8)
    assert len(got) == len(items_jpeg)
    for g, w_ in zip(got[:len(want)], want):
        assert "error" not in g, g
        for k in w_:
            assert g[k] == w_[k], (w_["path"], k)
    assert "error" in got[-2]                                  # corrupt entropy data
    assert bp.metrics.get("host_decoded") == 1                # progressive without a frame: read by the host loader (Pillow)
    from facet_b200.utils.image_loading import decode_on_host
    ref = bp.process_items_streamed([{"path": "/x/prog_only.jpg", "img_cv": decode_on_host(buf.getvalue())}])[0]
    assert "error" not in got[-1] and all(got[-1][k] == ref[k] for k in ref)


def test_streamed_side_products_leading_lines_and_thumbnails(scorer):
    """With `leading_lines` every result carries the score OpenCV's own calls give on the same frame
    (composition.py:190-261 restated with cv2 here), with `thumbnails` the JPEG bytes Pillow's
    `thumbnail((640, 640), LANCZOS)` + `save(quality=80)` give (scorer.py:1681-1686); the blocking path agrees."""
    import io
    import cv2
    from PIL import Image
    from facet_b200.analyzers.composition import CompositionAnalyzer
    from facet_b200.processing.batch_processor import BatchProcessor
    sc, tags, names = scorer
    shapes = [(700, 1050)] * 5 + [(1024, 683)] * 2 + [(97, 131)]
    items = [{"path": f"/x/c{i}.jpg", "img_cv": synth_image_bgr(i, h, w)} for i, (h, w) in enumerate(shapes)]
    items[3] = dict(items[3], leading_lines_score=4.25)               # supplied by the caller: kept as is
    items.insert(2, {"path": "/x/broken.jpg", "error": "Failed to load image"})
    bp = BatchProcessor(sc, batch_size=16, num_workers=3, leading_lines=True)
    got = bp.process_items_streamed(items, chunk=3, vit_batch=4, thumbnails=True)
    blocking = list(bp.process_items(items))
    plain = BatchProcessor(sc, batch_size=16).process_items_streamed(items, chunk=3, vit_batch=4)
    for it, g, b, p in zip(items, got, blocking, plain):
        if "error" in it:
            assert "error" in g
            continue
        img = it["img_cv"]
        h, w = img.shape[:2]
        if "leading_lines_score" in it:
            want = it["leading_lines_score"]
        else:
            edges = cv2.Canny(cv2.GaussianBlur(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (5, 5), 0), 50, 150)
            want = CompositionAnalyzer.score_lines(CompositionAnalyzer.lines_from_edges(edges, h, w), h, w)["leading_lines_score"]
        assert g["leading_lines_score"] == want == b["leading_lines_score"], (it["path"], g["leading_lines_score"], want)
        thumb = Image.fromarray(img[:, :, ::-1].copy())
        thumb.thumbnail((640, 640), Image.Resampling.LANCZOS)
        buf = io.BytesIO()
        thumb.save(buf, format="JPEG", quality=80)
        assert g["thumbnail"] == buf.getvalue(), it["path"]
        for k in p:                                                    # every other column as without the side products
            if k != "leading_lines_score":
                assert g[k] == p[k], (it["path"], k)
        assert "thumbnail" not in p


def test_score_photo_from_pil_single_image_engine(tmp_path):
    """`Facet.score_photo_from_pil` (scorer.py:952-1147): the keys of the reference's row, the analyzer columns of the batch
    pass, the composition values the (golden-tested) CompositionAnalyzer gives with the subject search switched on, and the
    aggregate of `calculate_aggregate_logic` on the single-image metric set (with `is_monochrome` / `contrast_score`)."""
    import json
    import os
    from PIL import Image
    from facet_b200.analyzers.composition import CompositionAnalyzer
    from facet_b200.config import ScoringConfig
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.processing.aggregate import calculate_aggregate_logic
    from facet_b200.processing.scorer import Facet
    with open(os.path.join(os.path.dirname(__file__), "golden", "aggregate_golden.json")) as f:
        cfg_dict = json.load(f)["cases"][0]["config"]
    path = tmp_path / "scoring_config.json"
    path.write_text(json.dumps(cfg_dict))
    cfg = ScoringConfig(str(path))
    sc = Facet(random_state_dict(0), config=cfg)
    img = synth_image_bgr(4, 683, 1024)
    pil = Image.fromarray(np.ascontiguousarray(img[..., ::-1]))
    photo = tmp_path / "sub" / ".." / "photo_0001.jpg"
    res = sc.score_photo_from_pil(pil, img, str(photo))
    reference_keys = {
        "path", "filename", "category", "image_width", "image_height", "aesthetic", "face_count", "face_quality", "eye_sharpness",
        "face_sharpness", "face_ratio", "tech_sharpness", "color_score", "exposure_score", "comp_score", "isolation_bonus", "is_blink",
        "phash", "aggregate", "clip_embedding", "raw_sharpness_variance", "histogram_data", "histogram_spread", "mean_luminance",
        "histogram_bimodality", "power_point_score", "raw_color_entropy", "raw_eye_sharpness", "config_version", "shadow_clipped",
        "highlight_clipped", "is_silhouette", "is_group_portrait", "leading_lines_score", "face_confidence", "is_monochrome",
        "mean_saturation", "dynamic_range_stops", "noise_sigma", "contrast_score", "tags", "quality_score", "topiq_score",
        "composition_explanation", "scoring_model", "composition_pattern", "face_details"}
    assert set(res) == reference_keys
    assert res["path"] == str((tmp_path / "photo_0001.jpg").resolve()) and res["filename"] == "photo_0001.jpg"
    d = sc.score_images(img[None], mono_threshold=cfg.get_monochrome_settings()["saturation_threshold_percent"] / 100)[0]
    for k in ("aesthetic", "tech_sharpness", "color_score", "exposure_score", "phash", "clip_embedding", "histogram_data", "noise_sigma",
              "contrast_score", "dynamic_range_stops", "is_monochrome", "mean_saturation", "raw_sharpness_variance"):
        assert res[k] == d[k], k
    comp = CompositionAnalyzer.get_placement_data(None, 1024, 683, cfg, img)          # subject search on the frame
    assert comp != {"score": 7.0, "power_point_score": 5.0, "line_score": 5.0, "center_score": 7.0}
    assert res["comp_score"] == round(comp["score"], 2) and res["power_point_score"] == float(comp["power_point_score"])
    assert res["leading_lines_score"] == CompositionAnalyzer.detect_leading_lines(img)["leading_lines_score"]
    agg, cat = calculate_aggregate_logic({
        "aesthetic": d["aesthetic_unrounded"], "face_count": 0, "face_quality": 0, "eye_sharpness": 0,
        "tech_sharpness": d["tech_sharpness_unrounded"], "color_score": d["color_score_unrounded"],
        "exposure_score": d["exposure_score_unrounded"], "face_ratio": 0.0, "comp_score": comp["score"], "isolation_bonus": 1.0,
        "is_blink": 0, "shadow_clipped": d["shadow_clipped"], "highlight_clipped": d["highlight_clipped"], "is_silhouette": 0,
        "histogram_spread": d["histogram_spread"], "is_monochrome": d["is_monochrome"], "contrast_score": d["contrast_score"],
        "iso": None, "f_stop": None}, cfg)
    assert res["aggregate"] == round(agg, 2) and res["category"] == cat
    assert sc.score_photo_from_pil(pil, np.zeros((1, 1, 3), np.uint8), str(photo)) is None        # failures print and return None
