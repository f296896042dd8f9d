"""D2H rate of one GPU with 1 / 2 / 4 concurrent copy streams (development tool): the link, not the stream count, is the limit
(55-57 GB/s on this pool whatever the number of streams or the chunk size)."""
import torch, time
n = 1600*1024*1024 // 8
src = torch.zeros(n, dtype=torch.float64, device='cuda')
dst = torch.empty(n, dtype=torch.float64).pin_memory()
def run(k, chunk_mb=256):
    streams = [torch.cuda.Stream() for _ in range(k)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step = chunk_mb * 1024 * 1024 // 8
    i = 0
    for off in range(0, n, step):
        s = streams[i % k]; i += 1
        with torch.cuda.stream(s):
            dst[off:off+step].copy_(src[off:off+step], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * 8 / dt / 1e9
for k in (1, 2, 4, 1, 2):
    print(k, 'streams', ' '.join(f"{run(k):.1f}" for _ in range(3)), 'GB/s', flush=True)
print('chunk 64MB, 2 streams', run(2, 64), 'chunk 1024MB 1 stream', run(1, 1024))
