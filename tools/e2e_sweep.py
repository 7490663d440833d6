import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from speech_emotion_privacy_trust_b200 import extraction, _lib
dev = torch.device('cuda', 0)
_lib.check(_lib.lib().sept_init(128))
lengths = bench.corpus_lengths(5531, 1234)
off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
wav = bench.synth_corpus_device(lengths, 4321, dev)
host_wav = torch.empty(wav.numel(), dtype=torch.float32, pin_memory=True); host_wav.copy_(wav)
frames = int(sum(1 + n // 160 for n in lengths))
host_out = torch.empty((frames, 128), dtype=torch.float32, pin_memory=True)
hours = lengths.sum() / 16000 / 3600
for cs in (1 << 25, 1 << 24, 1 << 23, 1 << 22):
    for ns in (3, 4, 6):
        for _ in range(2):
            extraction.logmel_host(host_wav, off, out_host=host_out, device=dev, chunk_samples=cs, n_streams=ns)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            extraction.logmel_host(host_wav, off, out_host=host_out, device=dev, chunk_samples=cs, n_streams=ns)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(f"chunk 2^{int(np.log2(cs))} streams {ns}: {ms:.2f} ms  {hours / ms * 1e3:.1f} audio-h/s")
