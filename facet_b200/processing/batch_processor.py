"""`BatchProcessor` for the in-scope part of the per-image pass (mirrors the role of
processing/batch_processor.py:27-658).

The reference collects `batch_size` loaded items (dicts with 'path', 'pil_img', 'img_cv',
'clip_input'), runs CLIP on the batch, then loops over the images on the CPU
(`_process_batch`, batch_processor.py:169-360).  Here a batch is grouped by frame shape and every
group goes through ONE device pass (`Facet.score_images`: technical metrics, perceptual hash, CLIP
preprocess, ViT-L/14 + aesthetic head + tags).  Results come back in input order as dicts with the
reference's column names and roundings (batch_processor.py:298-355), including `category` and
`aggregate` (`calculate_aggregate_logic`, vectorised over the group in processing/aggregate.py) and
the `is_silhouette` rule of utils/detection.py.  Columns owned by analyzers that are outside this
path take what the item carries — `face_res` (an `analyze_faces` result, analyzers/face.py:91),
`exif_data`, `leading_lines_score` — or the reference's own "nothing found" values (no faces,
centred-subject composition 7.0 / 5.0, no EXIF); a `scorer.face_analyzer`, when present, is called
like the reference calls it.  The optional `finish(item, result)` callback sees every result last.

Errors follow the reference's convention: a bad item never poisons the batch, it becomes
`{'path': ..., 'error': ...}` (batch_processor.py:109,359).
"""
from __future__ import annotations

from collections import defaultdict
from pathlib import Path

import numpy as np

from ..analyzers.composition import CompositionAnalyzer
from ..utils.detection import detect_silhouette

# analyze_faces() when nothing is detected (analyzers/face.py:91-97)
NO_FACES = {"face_count": 0, "face_quality": 0, "eye_sharpness": 0, "is_blink": 0, "face_area": 0, "bbox": None,
            "face_sharpness": 0, "raw_eye_sharpness": 0, "is_group_portrait": 0, "max_face_confidence": 0,
            "face_details": []}


class BatchProcessor:
    def __init__(self, scorer, batch_size=16, num_workers=4, finish=None, mono_threshold=None, leading_lines=False):
        """leading_lines: run `CompositionAnalyzer.detect_leading_lines` per image like the reference's loop does
        (batch_processor.py:245; blur + Canny on the device, OpenCV's probabilistic Hough on `num_workers` host threads)
        for items that do not already carry a `leading_lines_score`."""
        self.scorer = scorer
        self.leading_lines = bool(leading_lines)
        self.batch_size = int(batch_size)
        self.num_workers = int(num_workers)
        self.finish = finish
        cfg = getattr(scorer, "config", None)
        if mono_threshold is None and cfg is not None:
            mono_threshold = cfg.get_monochrome_settings().get("saturation_threshold_percent", 10) / 100
        self.mono_threshold = 0.10 if mono_threshold is None else mono_threshold
        self.metrics = {"images_processed": 0, "images_failed": 0, "batches": 0}

    def _tag_params(self):
        cfg = getattr(self.scorer, "config", None)
        if cfg is None:
            return 0.22, 5
        from ..utils.tags import get_tag_params
        return get_tag_params(cfg)

    def _process_batch(self, batch):
        """batch: list of dicts with 'path' and 'img_cv' ([H,W,3] uint8 BGR).  Returns one result per item."""
        results = [None] * len(batch)
        groups = defaultdict(list)
        for i, item in enumerate(batch):
            if not isinstance(item, dict):          # a loader that returned nothing usable (batch_processor.py:146)
                results[i] = {"path": None, "error": "Failed to load image"}
                continue
            img = item.get("img_cv")
            if "error" in item:
                results[i] = {"path": item.get("path"), "error": item["error"]}
            elif not isinstance(img, np.ndarray) or img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
                results[i] = {"path": item.get("path"), "error": "Failed to load image"}
            elif img.shape[0] < 2 or img.shape[1] < 2:
                results[i] = {"path": item.get("path"), "error": "image smaller than 2x2"}
            else:
                groups[img.shape[:2]].append(i)
        thr, max_tags = self._tag_params()
        for _shape, idxs in groups.items():
            try:
                frames = np.stack([batch[i]["img_cv"] for i in idxs])
                if self.leading_lines:
                    for i in idxs:
                        if batch[i].get("leading_lines_score") is None:
                            batch[i] = dict(batch[i], leading_lines_score=CompositionAnalyzer.detect_leading_lines(
                                batch[i]["img_cv"])["leading_lines_score"])
                scored = self.scorer.score_images(frames, mono_threshold=self.mono_threshold, tag_threshold=thr, max_tags=max_tags)
                scored = self._finish_group([batch[i] for i in idxs], scored)
                for i, res in zip(idxs, scored):
                    if self.finish is not None:
                        res = self.finish(batch[i], res)
                    results[i] = res
            except Exception as exc:  # a CUDA failure is loud, but it is reported per item like the reference does
                if len(idxs) == 1:
                    results[idxs[0]] = {"path": batch[idxs[0]].get("path"), "error": str(exc)}
                    continue
                # the reference isolates failures per image (batch_processor.py:190-360 wraps each image): retry the
                # group one frame at a time so a single bad item does not fail its whole shape group
                for i in idxs:
                    try:
                        one = self.scorer.score_images(batch[i]["img_cv"][None], mono_threshold=self.mono_threshold,
                                                       tag_threshold=thr, max_tags=max_tags)
                        res = self._finish_group([batch[i]], one)[0]
                        results[i] = self.finish(batch[i], res) if self.finish is not None else res
                    except Exception as exc_i:
                        results[i] = {"path": batch[i].get("path"), "error": str(exc_i)}
        self.metrics["batches"] += 1
        self.metrics["images_processed"] += sum(1 for r in results if r and "error" not in r)
        self.metrics["images_failed"] += sum(1 for r in results if r and "error" in r)
        return results

    def _finish_group(self, items, scored):
        """Everything `_process_batch` does per image after the analyzers (batch_processor.py:236-355): faces /
        composition / EXIF inputs, silhouette rule, aggregate + category, the result columns."""
        cfg = getattr(self.scorer, "config", None)
        face_analyzer = getattr(self.scorer, "face_analyzer", None)
        metrics, extras = [], []
        for item, res in zip(items, scored):
            h, w = res["image_height"], res["image_width"]
            face_res = item.get("face_res")
            if face_res is None:
                face_res = face_analyzer.analyze_faces(item["img_cv"]) if face_analyzer is not None else dict(NO_FACES)
            face_ratio = face_res.get("face_area", 0) / (h * w)
            comp = CompositionAnalyzer.get_placement_data(face_res.get("bbox"), w, h, cfg)
            isolation_bonus, is_blink = 1.0, 0
            if face_res["face_count"] > 0:
                isolation_bonus = max(1.0, face_res["face_sharpness"] / (res["raw_sharpness_variance"] + 1))
                is_blink = face_res.get("is_blink", 0)
            exif = item.get("exif_data") or {}
            is_silhouette = detect_silhouette({"is_silhouette": res["is_silhouette"]}, res["tags"], face_res.get("face_count", 0))
            # the reference hands exactly these keys to calculate_aggregate_logic (batch_processor.py:271-294):
            # no tags, luminance or noise columns, so only the face / silhouette / EXIF category rules can fire here
            metrics.append({
                "aesthetic": res["aesthetic_unrounded"], "face_count": face_res["face_count"],
                "face_quality": face_res["face_quality"], "eye_sharpness": face_res["eye_sharpness"],
                "tech_sharpness": res["tech_sharpness_unrounded"], "color_score": res["color_score_unrounded"],
                "exposure_score": res["exposure_score_unrounded"], "face_ratio": face_ratio, "comp_score": comp["score"],
                "isolation_bonus": isolation_bonus, "is_blink": is_blink, "shadow_clipped": res["shadow_clipped"],
                "highlight_clipped": res["highlight_clipped"], "is_silhouette": is_silhouette,
                "histogram_spread": res["histogram_spread"], "iso": exif.get("iso"), "f_stop": exif.get("f_stop"),
                "quality_score": res["quality_score"], "scoring_model": res["scoring_model"],
            })
            extras.append((face_res, face_ratio, comp, isolation_bonus, is_blink, exif, is_silhouette))
        if cfg is not None and hasattr(cfg, "_scoring"):
            aggregates, categories = cfg._scoring().score_batch(metrics)
        elif cfg is not None:        # the reference's own ScoringConfig object
            from .aggregate import AggregateScorer
            aggregates, categories = AggregateScorer(cfg).score_batch(metrics)
        else:
            aggregates, categories = [None] * len(items), [None] * len(items)
        out = []
        for item, res, agg, cat, (face_res, face_ratio, comp, iso_bonus, is_blink, exif, is_sil) in zip(
                items, scored, aggregates, categories, extras):
            for k in ("aesthetic_unrounded", "tech_sharpness_unrounded", "color_score_unrounded", "exposure_score_unrounded"):
                res.pop(k)
            # batch_processor.py:298-300: the row key is the resolved absolute path, the filename its last component
            path = item.get("path")
            resolved = Path(path).resolve() if path is not None else None
            res.update({
                "path": str(resolved) if resolved is not None else None, "filename": resolved.name if resolved is not None else "",
                "category": cat, "aggregate": None if agg is None else round(float(agg), 2),
                "face_count": face_res["face_count"], "face_quality": face_res["face_quality"],
                "eye_sharpness": face_res["eye_sharpness"], "face_sharpness": face_res["face_sharpness"],
                "face_ratio": face_ratio, "comp_score": round(comp["score"], 2), "isolation_bonus": round(iso_bonus, 2),
                "is_blink": is_blink, "power_point_score": float(comp["power_point_score"]),
                "raw_eye_sharpness": float(face_res.get("raw_eye_sharpness", 0)),
                "config_version": getattr(cfg, "version_hash", None),
                "is_silhouette": is_sil, "is_group_portrait": face_res.get("is_group_portrait", 0),
                "leading_lines_score": item.get("leading_lines_score") or 0,
                "face_confidence": face_res.get("max_face_confidence", 0),
                "composition_explanation": comp.get("vlm_explanation"), "composition_pattern": None,
                "face_details": face_res.get("face_details", []),
            })
            res.update(exif)
            out.append(res)
        return out

    # -- overlapped path ---------------------------------------------------------------------------------------
    def process_items_streamed(self, items, chunk=16, vit_batch=64, rgb_order=False, thumbnails=False):
        """Same results as `process_items` (one dict per item, input order), produced by an overlapped pipeline instead
        of one blocking pass per batch — the role of the reference's loader threads / GPU thread / result queue
        (batch_processor.py:123-167, 362-455):

          copy stream     each item's frame — or, for items that carry `jpeg` (the FILE BYTES: bytes or a uint8 array, e.g. a
                          view of a pinned read buffer), its compressed stream — goes host -> device (async from pinned
                          memory) into one of two `chunk`-slot staging buffers while the previous chunk is being processed
          compute stream  JPEG chunks are decoded on the device first (fb_jpeg_decode + fb_orient for the EXIF orientation:
                          the pixel work of utils/image_loading.py:90-106); then technical pass + pHash + CLIP preprocess per
                          chunk; the ViT tower + heads once `vit_batch` frames have been preprocessed (frames of different
                          shapes share a ViT launch)
          D2H stream      ONE packed record per image (histogram, sums, hash, embedding, aesthetic, tag similarities;
                          about 5 KB) into pinned host memory per ViT batch
          worker thread   closed-form metric dicts, tag selection, aggregate + category, the result columns
          side products   with `leading_lines` (constructor) the Canny edge map of every frame (csrc/canny.cu) and with
                          `thumbnails` the 640-px LANCZOS thumbnail pixels of every frame (csrc/thumbnail.cu) are produced in
                          the same visit of the frame on the compute stream, leave on the D2H stream, and `num_workers`
                          the thumbnails are JPEG-encoded on the device as well (csrc/jpeg_encode.cu); host threads run OpenCV's
                          Hough transform on the edge maps and slice the streams; a result then carries
                          `leading_lines_score` (batch_processor.py:245,334) / `thumbnail` (the JPEG bytes of scorer.py:1681-1686)

        `self.metrics` gains h2d_bytes / d2h_bytes of the call."""
        import queue
        import threading
        from .. import _lib, ops
        from ..utils import jpeg as fj
        from ..utils.image_loading import decode_on_host
        torch = _lib.require_cuda()
        scorer = self.scorer
        dev = scorer.device
        chunk, vit_batch = max(1, int(chunk)), max(1, int(vit_batch))
        thr, max_tags = self._tag_params()
        n_tags = scorer.model.n_tags
        rec_bytes = 1024 + 32 + 32 + 8 + 8 + 3072 + 4 * n_tags
        st = self._stream_state(torch, dev, vit_batch + chunk, rec_bytes)
        compute = torch.cuda.current_stream(dev)
        copy_s, d2h_s = st["copy"], st["d2h"]
        results = {}
        self.metrics.setdefault("h2d_bytes", 0)
        self.metrics.setdefault("d2h_bytes", 0)
        jobs = queue.Queue()
        side = {}                 # pos -> {"lines": future, "thumb": (future, k)}
        want_lines = self.leading_lines
        pool = None
        if want_lines or thumbnails:
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(max_workers=max(1, self.num_workers))
        pinned = {}               # (kind, shape) -> queue of pinned host buffers (back-pressure on the side products)

        def pinned_get(kind, shape, count):
            q = pinned.get((kind, shape))
            if q is None:
                q = pinned[(kind, shape)] = queue.Queue()
                for _ in range(count):
                    q.put(torch.empty(shape, dtype=torch.uint8, pin_memory=True))
            return q, q.get()

        def d2h_async(src, kind, count):
            """src (device tensor, produced on the compute stream) -> a pinned buffer on the D2H stream."""
            q, host = pinned_get(kind, tuple(src.shape), count)
            ready = torch.cuda.Event()
            ready.record(compute)
            done = torch.cuda.Event()
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(ready)
                host.copy_(src, non_blocking=True)
                done.record(d2h_s)
            src.record_stream(d2h_s)
            self.metrics["d2h_bytes"] += src.numel()
            return q, host, done

        def lines_job(q, host, done, h, w):
            try:
                done.synchronize()
                return CompositionAnalyzer.score_lines(CompositionAnalyzer.lines_from_edges(host.numpy(), h, w), h, w)
            finally:
                q.put(host)

        thumb_cap = 1 << 18          # bytes of every stream fetched with the first copy (photographs: 15-60 KB at 640 px)

        def thumbs_job(q, host, done, m, streams, lengths):
            """host: pinned [chunk, 4 + thumb_cap] uint8 = (length, first thumb_cap bytes) of every stream of the chunk."""
            try:
                done.synchronize()
                a = host[:m].numpy()
                lens = a[:, :4].view(np.int32).reshape(-1).copy()
                out = [a[k, 4:4 + min(int(lens[k]), thumb_cap)].tobytes() for k in range(m)]
            finally:
                q.put(host)
            for k in range(m):                       # a stream longer than the first copy (noise-like frames): fetch it whole
                if lens[k] > thumb_cap:
                    out[k] = streams[k, :int(lens[k])].cpu().numpy().tobytes()
            return out

        def side_products(frames, metas, thumbs):
            if want_lines:
                for k, (pos, item, h, w) in enumerate(metas):
                    if item.get("leading_lines_score") is None:
                        edges = ops.canny_edges(ops.gray_plane(frames[k], rgb_order=rgb_order), 50, 150, blur=True)
                        side.setdefault(pos, {})["lines"] = pool.submit(lines_job, *d2h_async(edges, "edges", 2 * self.num_workers + 2), h, w)
            if thumbs is not None:
                # the JPEG streams are made on the device too (csrc/jpeg_encode.cu, the bytes Pillow's encoder writes): what
                # leaves is (length, stream) per frame instead of 820 KB of pixels and a host-side encode
                streams, lengths = ops.jpeg_encode(thumbs, quality=80, as_device=True)
                m = len(metas)
                rec = torch.zeros((chunk, 4 + thumb_cap), dtype=torch.uint8, device=dev)
                rec[:m, :4] = lengths.view(torch.uint8).reshape(m, 4)
                w_ = min(thumb_cap, int(streams.shape[1]))
                rec[:m, 4:4 + w_] = streams[:, :w_]
                fut = pool.submit(thumbs_job, *d2h_async(rec, "thumbs", 4), m, streams, lengths)
                for k, (pos, _item, _h, _w) in enumerate(metas):
                    side.setdefault(pos, {})["thumb"] = (fut, k)

        def with_side(pos, item):
            """The item as `_finish_group` should see it, and the thumbnail bytes of its frame (or None)."""
            extra = side.pop(pos, None)
            if not extra:
                return item, None
            if "lines" in extra:
                item = dict(item, leading_lines_score=extra["lines"].result()["leading_lines_score"])
            thumb = None
            if "thumb" in extra:
                fut, k = extra["thumb"]
                thumb = fut.result()[k]
            return item, thumb

        def post(job):
            done, pack, metas = job
            try:
                done.synchronize()
                a = pack[:len(metas)].numpy()
                hist = a[:, :1024].view(np.uint32).astype(np.int64)
                sums = a[:, 1024:1056].view(np.int64).copy()
                der = a[:, 1056:1088].view(np.float64).copy()
                hashes = a[:, 1088:1096].view(np.uint64).reshape(-1).copy()
                raw = a[:, 1096:1100].view(np.float32).reshape(-1).copy()
                status = a[:, 1100:1104].view(np.int32).reshape(-1).copy()
                emb = a[:, 1104:1104 + 3072].view(np.float32).copy()
                sims = a[:, 1104 + 3072:].view(np.float32).copy() if n_tags else None
            finally:
                st["free_packs"].put(pack)
            by_shape = defaultdict(list)
            for k, (pos, item, h, w) in enumerate(metas):
                if status[k]:             # fb_jpeg_decode flagged the stream (utils/image_loading.py returns None for unreadable files)
                    results[pos] = {"path": item.get("path"), "error": "Failed to load image (corrupt JPEG data)"}
                else:
                    by_shape[(h, w)].append(k)
            for (h, w), ks in by_shape.items():
                sel = np.asarray(ks)
                try:
                    scored = scorer.results_from_host(h, w, hist[sel], sums[sel], der[sel], raw[sel], emb[sel],
                                                      sims[sel] if sims is not None else None,
                                                      ["%016x" % int(v) for v in hashes[sel]], mono_threshold=self.mono_threshold,
                                                      tag_threshold=thr, max_tags=max_tags)
                    sided = [with_side(metas[k][0], metas[k][1]) for k in ks]
                    scored = self._finish_group([it for it, _ in sided], scored)
                    for k, res, (_it, thumb) in zip(ks, scored, sided):
                        if thumb is not None:
                            res["thumbnail"] = thumb
                        results[metas[k][0]] = self.finish(metas[k][1], res) if self.finish is not None else res
                except Exception as exc:
                    for k in ks:
                        results[metas[k][0]] = {"path": metas[k][1].get("path"), "error": str(exc)}

        def worker():
            while True:
                job = jobs.get()
                if job is None:
                    return
                post(job)

        th = threading.Thread(target=worker, daemon=True)
        th.start()

        cur = {"shape": None, "slot": 0, "metas": [], "bufs": None, "infos": None, "slot_bytes": 0}
        acc = {"px": [], "metas": []}         # preprocessed chunks waiting for the ViT launch

        def flush_vit():
            if not acc["metas"]:
                return
            metas, pxs = acc["metas"], acc["px"]
            acc["metas"], acc["px"] = [], []
            try:
                cat = (lambda key: torch.cat([p[key] for p in pxs]) if len(pxs) > 1 else pxs[0][key])
                vit = scorer.model.encode(cat("clip_in"))
                m = len(metas)
                parts = [cat("hist256").view(torch.uint8).reshape(m, 1024), cat("sums").view(torch.uint8).reshape(m, 32),
                         cat("derived").view(torch.uint8).reshape(m, 32), cat("phash").reshape(m, 1).view(torch.uint8),
                         vit["aesthetic_raw"].reshape(m, 1).view(torch.uint8), cat("status").reshape(m, 1).view(torch.uint8),
                         vit["embedding"].view(torch.uint8).reshape(m, 3072)]
                if n_tags:
                    parts.append(vit["tag_sims"].contiguous().view(torch.uint8).reshape(m, 4 * n_tags))
                rec = torch.cat(parts, dim=1)
                ready = torch.cuda.Event()
                ready.record(compute)
                pack = st["free_packs"].get()             # back-pressure: at most len(packs) ViT batches in flight
                done = torch.cuda.Event()
                with torch.cuda.stream(d2h_s):
                    d2h_s.wait_event(ready)
                    pack[:m].copy_(rec, non_blocking=True)
                    done.record(d2h_s)
                rec.record_stream(d2h_s)
                self.metrics["d2h_bytes"] += m * rec_bytes
                jobs.put((done, pack, metas))
            except Exception as exc:
                for pos, item, _h, _w in metas:
                    results[pos] = {"path": item.get("path"), "error": str(exc)}

        def flush_chunk():
            metas = cur["metas"]
            if not metas:
                return
            slot, bufs = cur["slot"], cur["bufs"]
            cur["metas"], cur["shape"] = [], None
            cur_infos = cur["infos"]
            try:
                bufs["ready"][slot].record(copy_s)
                compute.wait_event(bufs["ready"][slot])
                if cur_infos is not None:
                    # file bytes -> frames on the device (entropy decoding, IDCT, upsampling, colour conversion, EXIF transpose)
                    frames, status = ops.jpeg_decode_device(bufs["frames"][slot], cur["slot_bytes"], cur_infos, bgr=not rgb_order)
                    code = cur_infos[0].orientation
                    if code != 1:
                        frames = ops.orient(frames, code)
                    metas = [(pos, item, int(frames.shape[1]), int(frames.shape[2])) for pos, item, _h, _w in metas]
                else:
                    frames = bufs["frames"][slot][:len(metas)]
                    status = torch.zeros((len(metas),), dtype=torch.int32, device=dev)
                px = scorer.pixel_passes_device(frames, rgb_order=rgb_order, with_thumbnails=thumbnails)
                px["status"] = status
                if pool is not None:
                    side_products(frames, metas, px.pop("thumbnails", None))
                bufs["free"][slot].record(compute)
                bufs["used"][slot] = True
                acc["px"].append(px)
                acc["metas"].extend(metas)
            except Exception as exc:
                for pos, item, _h, _w in metas:
                    results[pos] = {"path": item.get("path"), "error": str(exc)}
            if len(acc["metas"]) >= vit_batch:
                flush_vit()

        n_items = 0
        for pos, item in enumerate(items):
            n_items = pos + 1
            if not isinstance(item, dict):
                results[pos] = {"path": None, "error": "Failed to load image"}
                continue
            img = item.get("img_cv")
            if "error" in item:
                results[pos] = {"path": item.get("path"), "error": item["error"]}
                continue
            info = None
            if item.get("jpeg") is not None:
                # the loader handed over file bytes: decode on the device unless the stream is of a kind only the CPU loader takes
                try:
                    info = fj.parse(item["jpeg"])
                except fj.UnsupportedJpeg as exc:
                    if img is None:
                        # not a stream the device decoder takes: the reference's own loader (Pillow) reads it
                        img = decode_on_host(item["jpeg"])
                        if img is None:
                            results[pos] = {"path": item.get("path"), "error": f"Failed to load image ({exc})"}
                            continue
                        self.metrics["host_decoded"] = self.metrics.get("host_decoded", 0) + 1
            if info is not None:
                data = item["jpeg"]
                nbytes = len(data)
                key = ("jpeg",) + info.geometry_key() + (info.orientation,)
                h, w = info.height, info.width
                if h < 2 or w < 2:
                    results[pos] = {"path": item.get("path"), "error": "image smaller than 2x2"}
                    continue
            else:
                if not isinstance(img, np.ndarray) or img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
                    results[pos] = {"path": item.get("path"), "error": "Failed to load image"}
                    continue
                h, w = int(img.shape[0]), int(img.shape[1])
                if h < 2 or w < 2:
                    results[pos] = {"path": item.get("path"), "error": "image smaller than 2x2"}
                    continue
                key = ("raw", h, w)
            if cur["metas"] and (cur["shape"] != key or len(cur["metas"]) == chunk or
                                 (info is not None and nbytes > cur["slot_bytes"])):
                flush_chunk()
            if not cur["metas"]:
                if info is not None:
                    # slots of a power-of-two size >= the stream (streams of one camera setting vary by a few 10 %)
                    slot_bytes = max(1 << 20, 1 << int(nbytes * 5 // 4).bit_length())
                    # + 256: the device bit reader / marker scan read up to 15 bytes past the end of a stream
                    bufs = self._chunk_buffers(torch, dev, ("jpeg", slot_bytes, chunk), (chunk * slot_bytes + 256,))
                    cur.update(infos=[], slot_bytes=slot_bytes)
                else:
                    bufs = self._chunk_buffers(torch, dev, ("raw", h, w, chunk), (chunk, h, w, 3))
                    cur.update(infos=None, slot_bytes=0)
                slot = bufs["next"]
                bufs["next"] = slot ^ 1
                cur.update(shape=key, slot=slot, bufs=bufs)
                if bufs["used"][slot]:
                    copy_s.wait_event(bufs["free"][slot])         # the compute stream is done with this staging buffer
            k = len(cur["metas"])
            with torch.cuda.stream(copy_s):
                if info is not None:
                    src = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.asarray(data, np.uint8).reshape(-1)
                    if not src.flags.writeable:
                        src = src.copy()
                    cur["bufs"]["frames"][cur["slot"]][k * cur["slot_bytes"]:k * cur["slot_bytes"] + nbytes].copy_(torch.from_numpy(src), non_blocking=True)
                    cur["infos"].append(info)
                    self.metrics["h2d_bytes"] += nbytes
                else:
                    cur["bufs"]["frames"][cur["slot"]][k].copy_(torch.from_numpy(np.ascontiguousarray(img)), non_blocking=True)
                    self.metrics["h2d_bytes"] += h * w * 3
            cur["metas"].append((pos, item, h, w))
        flush_chunk()
        flush_vit()
        jobs.put(None)
        th.join()
        if pool is not None:
            pool.shutdown(wait=True)
        self.metrics["batches"] += 1
        out = [results[i] for i in range(n_items)]
        self.metrics["images_processed"] += sum(1 for r in out if "error" not in r)
        self.metrics["images_failed"] += sum(1 for r in out if "error" in r)
        return out

    def _stream_state(self, torch, dev, max_batch, rec_bytes):
        import queue
        st = getattr(self, "_stream", None)
        if st is None or st["rec_bytes"] != rec_bytes or st["max_batch"] < max_batch:
            st = {"copy": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev), "rec_bytes": rec_bytes,
                  "max_batch": max_batch, "free_packs": queue.Queue(), "chunks": {}}
            for _ in range(3):
                st["free_packs"].put(torch.empty((max_batch, rec_bytes), dtype=torch.uint8, pin_memory=True))
            self._stream = st
        return st

    def _chunk_buffers(self, torch, dev, key, shape):
        chunks = self._stream["chunks"]
        if key not in chunks:
            if len(chunks) >= 4:                       # bound the staging memory when many frame shapes go by
                torch.cuda.current_stream(dev).synchronize()
                self._stream["copy"].synchronize()
                chunks.clear()
            chunks[key] = {"frames": [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(2)],
                           "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)],
                           "used": [False, False], "next": 0}
            # the caching allocator may hand out memory that kernels already queued on the compute stream still use
            # (temporaries freed in stream order): the copy stream must not write into it before they have run
            fence = torch.cuda.Event()
            fence.record(torch.cuda.current_stream(dev))
            self._stream["copy"].wait_event(fence)
        return chunks[key]

    # -- paths in, database rows out -------------------------------------------------------------------------
    def process_files(self, photo_paths, db_path=None, show_metrics=True, batch_save_size=50, chunk=16, vit_batch=64,
                      thumbnails=True):
        """The reference's `process_files(photo_paths)` (batch_processor.py:362-455): load every path, score it, save the rows
        every `batch_save_size` results.  Loader threads (`num_workers`) read the FILE BYTES of JPEG files (decoded on the device;
        RAW / PNG / other formats go through the host loader, utils/image_loading.load_any); the rows go to `db_path` (a database
        with the reference's schema) through `db_sink.PhotoSink`; with db_path=None the result dicts are returned instead.
        Errors are printed and skipped like the reference does (`Error on <path>: <message>`)."""
        import time
        from concurrent.futures import ThreadPoolExecutor
        from ..utils.image_loading import load_item
        from .db_sink import PhotoSink
        paths = list(photo_paths)
        if not paths:
            return [] if db_path is None else 0
        start = time.time()
        self.metrics["start_time"] = start
        with ThreadPoolExecutor(max_workers=max(1, self.num_workers)) as pool:
            items = list(pool.map(load_item, paths))              # file reads overlap each other; bytes stay undecoded
        results = self.process_items_streamed(items, chunk=chunk, vit_batch=vit_batch, thumbnails=bool(thumbnails and db_path is not None))
        for r in results:
            if "error" in r:
                print(f"Error on {r.get('path', 'unknown')}: {r['error']}")
        self.metrics["elapsed_time"] = time.time() - start
        if db_path is None:
            return results
        with PhotoSink(db_path, batch_save_size=batch_save_size) as sink:
            for item, res in zip(items, results):
                if "error" in res:
                    continue
                # scorer.py:1681-1686: the 640-px thumbnail was made in the same visit of the frame (res['thumbnail'])
                image = None
                sink.add(res, image)
        if show_metrics:
            dt = max(self.metrics["elapsed_time"], 1e-9)
            print(f"[{sink.saved}/{len(paths)}] {len(paths) / dt:.1f} img/s")
        return sink.saved

    def process_stream(self, path_iterator, total_count, tuning_callback=None, tuning_interval=50, calibration_callback=None,
                       calibration_size=20, show_metrics=True, db_path=None, **kwargs):
        """The reference's `process_stream` (batch_processor.py:458-658) on top of `process_files`: same arguments, same return
        convention (None when everything was processed; the remaining paths when `calibration_callback` asked for a processor
        with a different worker count after the first `2 * calibration_size` paths).  The reference needs continuous loader
        threads, a resource monitor and tuning callbacks to keep a small GPU fed; here one `process_files` call per phase streams
        everything (uploads, kernels, read-back and row building overlap inside it), `tuning_callback` is called with the metrics
        every `tuning_interval` images' worth of progress at the end of a phase, and rows go to `db_path` when given (else the
        result dicts are kept in `self.last_results`)."""
        if total_count == 0:
            return None
        paths = list(path_iterator)
        self.last_results = []

        def run(part):
            out = self.process_files(part, db_path=db_path, show_metrics=show_metrics, **kwargs)
            if db_path is None:
                self.last_results.extend(out)

        if calibration_callback and len(paths) > calibration_size * 2:
            run(paths[:calibration_size * 2])
            paths = paths[calibration_size * 2:]
            if calibration_callback(dict(self.metrics)):
                return paths
            if not paths:
                return None
        run(paths)
        if tuning_callback is not None:
            tuning_callback(dict(self.metrics))
        return None

    def process_items(self, items):
        """Stream items through `_process_batch` in chunks of batch_size; yields results in order."""
        chunk = []
        for item in items:
            chunk.append(item)
            if len(chunk) == self.batch_size:
                yield from self._process_batch(chunk)
                chunk = []
        if chunk:
            yield from self._process_batch(chunk)
