"""Batched enhancement of utterances of DIFFERENT lengths (the reference never batches inference -
test_nsvae_se.py enhances one file per forward).

The causal network has no cross-utterance operation and, frame by frame, no dependence on later frames, so a
zero-padded batch reproduces the per-utterance results exactly as long as the two length-dependent steps see each
utterance's own length: the STFT's reflect padding at the END of the signal and the iSTFT's window-envelope / trim at
the end.  ``STFT.lengths`` / ``ISTFT.lengths`` (int32 (B,)) carry them into ``idv_stft_frames_split`` / ``idv_ola_fwd``.

``bucket_batches`` groups utterances of similar length so that little of a batch is padding; ``enhance_ragged`` runs the
buckets with pinned-memory host->device staging on a side stream (the copy of bucket i+1 overlaps the forward of
bucket i) and returns one waveform per input utterance, each of the length the reference returns for it alone,
``hop * (L // hop)``."""
import torch


def bucket_batches(lengths, max_batch=64, max_pad_frac=0.1, multiple=100):
    """Indices of ``lengths`` grouped into batches: sorted by length (longest first), a batch is closed when it holds
    ``max_batch`` utterances or when adding the next one would make the padded samples exceed ``max_pad_frac`` of the
    batch.  Returns [(indices, padded_length)], padded_length a multiple of ``multiple`` (the hop)."""
    if max_batch < 1 or not 0 <= max_pad_frac < 1:
        raise ValueError("max_batch >= 1 and 0 <= max_pad_frac < 1 expected")
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    out, cur, cur_len, cur_sum = [], [], 0, 0
    for i in order:
        L = int(lengths[i])
        if L <= 0:
            raise ValueError("utterance %d is empty" % i)
        if cur:
            waste = 1.0 - (cur_sum + L) / float(cur_len * (len(cur) + 1))
            if len(cur) >= max_batch or waste > max_pad_frac:
                out.append((cur, cur_len))
                cur, cur_sum = [], 0
        if not cur:
            cur_len = (L + multiple - 1) // multiple * multiple
        cur.append(i)
        cur_sum += L
    if cur:
        out.append((cur, cur_len))
    return out


def _set_lengths(encoder, decoder, lengths):
    encoder.stft.lengths = lengths
    decoder.istft.lengths = lengths
    if getattr(decoder, "stft", None) is not None:
        decoder.stft.lengths = lengths


def enhance_ragged(waves, encoder, decoder, device, max_batch=64, max_pad_frac=0.1, decoder_kwargs=None, eps_fn=None):
    """waves: list of 1-D float32 CPU tensors (16 kHz utterances of any lengths > n_fft/2).  Returns the list of
    enhanced waveforms (CPU tensors, utterance i has ``hop * (len_i // hop)`` samples, exactly what
    ``decoder(encoder(wave_i[None]))`` returns).  eps_fn(indices, padded_frames) -> supplied eps list or None."""
    if not getattr(encoder, "causal", False):
        raise NotImplementedError("ragged batching relies on causality (model/causal_netconfig.py)")
    if getattr(decoder, "num_samples", 1) != 1:
        raise NotImplementedError("ragged batching is built for num_samples = 1")
    hop = encoder.stft.hop_length
    decoder_kwargs = decoder_kwargs or {}
    lens = [int(w.numel()) for w in waves]
    n_fft = encoder.stft.n_fft
    short = [i for i, n in enumerate(lens) if n <= n_fft // 2]
    if short:          # torch.stft's reflect padding (model/pvae_module.py:L22) raises for these; so do we
        raise RuntimeError("utterance(s) %s have %s samples: the STFT's reflect padding needs more than n_fft/2 = %d"
                           % (short[:8], [lens[i] for i in short[:8]], n_fft // 2))
    buckets = bucket_batches(lens, max_batch, max_pad_frac, hop)
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
    results = [None] * len(waves)

    def stage(bucket):
        idx, L = bucket
        host = torch.zeros((len(idx), L), dtype=torch.float32)
        for r, i in enumerate(idx):
            host[r, :lens[i]] = waves[i]
        hl = torch.tensor([lens[i] for i in idx], dtype=torch.int32)
        if copy_stream is None:
            return host.to(dev), hl.to(dev), None
        host, hl = host.pin_memory(), hl.pin_memory()
        with torch.cuda.stream(copy_stream):
            x, l = host.to(dev, non_blocking=True), hl.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x, l, ev

    staged = stage(buckets[0]) if buckets else None
    try:
        for k, (idx, L) in enumerate(buckets):
            x, l, ev = staged
            staged = stage(buckets[k + 1]) if k + 1 < len(buckets) else None      # overlaps this bucket's forward
            if ev is not None:
                torch.cuda.current_stream(dev).wait_event(ev)
                x.record_stream(torch.cuda.current_stream(dev))
                l.record_stream(torch.cuda.current_stream(dev))
            _set_lengths(encoder, decoder, l)
            with torch.no_grad():
                eps = eps_fn(idx, L // hop + 1) if eps_fn is not None else None
                r = encoder(x, train=False, eps=eps) if eps is not None else encoder(x, train=False)
                stft_x, z, skiper, C, F = r[-1], r[0], r[-4], r[-3], r[-2]
                sig, _ = decoder(stft_x, z, skiper, C, F, train=False, **decoder_kwargs)
            sig = sig.cpu()
            for row, i in enumerate(idx):
                results[i] = sig[row, :hop * (lens[i] // hop)].clone()
    finally:
        _set_lengths(encoder, decoder, None)
    return results
