import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from office_person_detection_vit_b200.detection import ViTDetector
from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames
det = ViTDetector(confidence_threshold=0.5, state_dict=random_init_state_dict(0)); det.load_model()
frames = [f for f in synthetic_frames(64, 800, 1333, seed=1)]
det.detect_batch(frames)
for _ in range(2):
    t0 = time.perf_counter(); r = det.detect_batch(frames); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("detect_batch", (t1 - t0) * 1e3, "ms", sum(len(d) for d in r))
# pieces
stage = det._staging(64, 800, 1333); host = stage.numpy()
t0 = time.perf_counter()
for j, f in enumerate(frames): host[j] = f
t1 = time.perf_counter(); print("staging copy", (t1 - t0) * 1e3)
dev = torch.device("cuda")
t0 = time.perf_counter(); d = stage.to(dev, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter(); print("h2d", (t1 - t0) * 1e3)
t0 = time.perf_counter(); out = det.detect_tensors(d); torch.cuda.synchronize(); t1 = time.perf_counter(); print("forward+post", (t1 - t0) * 1e3)
t0 = time.perf_counter(); dets, _ = det._to_detections(out, None, 64); t1 = time.perf_counter(); print("to_detections", (t1 - t0) * 1e3)
import concurrent.futures as cf
pool = cf.ThreadPoolExecutor(8)
def cp(j): host[j] = frames[j]
t0 = time.perf_counter(); list(pool.map(cp, range(64))); t1 = time.perf_counter(); print("staging copy 8 threads", (t1 - t0) * 1e3)
