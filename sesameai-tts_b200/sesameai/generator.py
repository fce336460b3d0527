"""Drop-in ``sesameai.generator`` for B200.

Public surface of the reference module (``/root/reference/sesameai/generator.py``): ``Segment``,
``Generator`` (``generate``, ``generate_stream``, ``_tokenize_text_segment``, ``_tokenize_audio``,
``_tokenize_segment``, attributes ``sample_rate`` / ``device`` / ``_model`` / ``_audio_tokenizer`` /
``_text_tokenizer`` / ``_stream_buffer_size``), ``load_csm_1b`` and ``load_llama3_tokenizer`` -- so
``tts_service.py`` and the web front-ends run unchanged.  Frame generation goes through
``Model.generate_frame`` (one persistent CUDA kernel per frame) and audio through ``MimiCodec``.

Offline use: ``Generator(model, text_tokenizer=..., audio_tokenizer=...)`` accepts injected
tokenizers (the reference downloads both from the HF hub inside ``__init__``).
"""
from __future__ import annotations

import queue
import threading
import time
from dataclasses import dataclass
from typing import Callable, Iterator, List, Optional, Tuple

import torch

from .mimi import DEFAULT_REPO, MIMI_NAME, MimiCodec, get_mimi
from .models import Model

FRAME_MS = 80  # one Mimi frame at 12.5 Hz
MAX_SEQ_LEN = 2048


@dataclass
class Segment:
    speaker: int
    text: str
    audio: torch.Tensor  # (num_samples,), 24 kHz


def load_llama3_tokenizer():
    """Llama-3 tokenizer with the BOS/EOS template the reference installs (``generator.py:24-38``)."""
    from tokenizers.processors import TemplateProcessing
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained("meta-llama/Llama-3.2-1B")
    bos, eos = tok.bos_token, tok.eos_token
    tok._tokenizer.post_processor = TemplateProcessing(
        single=f"{bos}:0 $A:0 {eos}:0",
        pair=f"{bos}:0 $A:0 {eos}:0 {bos}:1 $B:1 {eos}:1",
        special_tokens=[(bos, tok.bos_token_id), (eos, tok.eos_token_id)],
    )
    return tok


class Generator:
    def __init__(self, model: Model, text_tokenizer=None, audio_tokenizer: Optional[MimiCodec] = None):
        self._model = model
        self._model.setup_caches(1)
        self.device = next(model.parameters()).device
        self._text_tokenizer = text_tokenizer if text_tokenizer is not None else load_llama3_tokenizer()
        if audio_tokenizer is None:
            from huggingface_hub import hf_hub_download

            audio_tokenizer = get_mimi(hf_hub_download(DEFAULT_REPO, MIMI_NAME), device=self.device)
        audio_tokenizer.set_num_codebooks(32)
        self._audio_tokenizer = audio_tokenizer
        self.sample_rate = audio_tokenizer.sample_rate
        self._stream_buffer_size = 10  # frames per streamed chunk
        # generate_stream decodes its chunks with a stateful Mimi stream: the causal left context (conv tails,
        # transformer K/V) is carried over, so the concatenated chunks equal the one-shot decode of generate().
        # False = the reference's behaviour (generator.py:111-117,189-196): every chunk decoded from silence.
        self._stream_stateful = True
        self._n_cols = model.config.audio_num_codebooks + 1

    # -- prompt frames ---------------------------------------------------------------------------
    def _tokenize_text_segment(self, text: str, speaker: int) -> Tuple[torch.Tensor, torch.Tensor]:
        ids = self._text_tokenizer.encode(f"[{speaker}]{text}")
        frame = torch.zeros(len(ids), self._n_cols, dtype=torch.long)
        mask = torch.zeros(len(ids), self._n_cols, dtype=torch.bool)
        frame[:, -1] = torch.tensor(ids, dtype=torch.long)
        mask[:, -1] = True
        return frame.to(self.device), mask.to(self.device)

    def _tokenize_audio(self, audio: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert audio.ndim == 1, "Audio must be single channel"
        codes = self._audio_tokenizer.encode(audio.to(self.device).unsqueeze(0).unsqueeze(0))[0]  # (K, T)
        codes = torch.cat([codes, torch.zeros(codes.size(0), 1, dtype=codes.dtype, device=codes.device)], dim=1)  # EOS frame
        frame = torch.zeros(codes.size(1), self._n_cols, dtype=torch.long, device=self.device)
        mask = torch.zeros(codes.size(1), self._n_cols, dtype=torch.bool, device=self.device)
        frame[:, :-1] = codes.transpose(0, 1)
        mask[:, :-1] = True
        return frame, mask

    def _tokenize_segment(self, segment: Segment) -> Tuple[torch.Tensor, torch.Tensor]:
        tt, tm = self._tokenize_text_segment(segment.text, segment.speaker)
        at, am = self._tokenize_audio(segment.audio)
        return torch.cat([tt, at], dim=0), torch.cat([tm, am], dim=0)

    def _build_prompt(self, text: str, speaker: int, context: List[Segment], max_generation_len: int):
        parts = [self._tokenize_segment(seg) for seg in context]
        parts.append(self._tokenize_text_segment(text, speaker))
        tokens = torch.cat([p[0] for p in parts], dim=0).long().to(self.device)
        mask = torch.cat([p[1] for p in parts], dim=0).bool().to(self.device)
        limit = MAX_SEQ_LEN - max_generation_len
        if tokens.size(0) >= limit:
            raise ValueError(f"Inputs too long, must be below max_seq_len - max_generation_len: {limit}")
        pos = torch.arange(0, tokens.size(0), device=self.device).unsqueeze(0).long()
        return tokens.unsqueeze(0), mask.unsqueeze(0), pos

    def _decode_frames(self, frames: List[torch.Tensor]) -> torch.Tensor:
        if not frames:
            return torch.tensor([])
        return self._audio_tokenizer.decode(torch.stack(frames).permute(1, 2, 0)).squeeze(0).squeeze(0)

    def _frames(self, text, speaker, context, max_audio_length_ms, temperature, topk) -> Iterator[torch.Tensor]:
        """The reference frame loop (``generator.py:283-294``): one ``generate_frame`` per 80 ms,
        stop at the all-zero EOS frame, feed the sampled codes back as the next input frame."""
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        self._model.reset_caches()
        max_generation_len = int(max_audio_length_ms / FRAME_MS)
        tokens, mask, pos = self._build_prompt(text, speaker, context, max_generation_len)
        nxt_tok = torch.zeros(1, 1, self._n_cols, dtype=torch.long, device=self.device)
        nxt_mask = torch.ones(1, 1, self._n_cols, dtype=torch.bool, device=self.device)
        nxt_mask[..., -1] = False
        for _ in range(max_generation_len):
            sample = self._model.generate_frame(tokens, mask, pos, temperature, topk)
            eos = bool(torch.all(sample == 0))  # synchronises the stream, like the reference's ``if``
            self._model.check_device_error()    # bad token ids / positions or a drained kernel raise here
            if eos:
                return  # EOS
            yield sample
            nxt_tok[0, 0, :-1] = sample[0]
            tokens, mask = nxt_tok, nxt_mask
            pos = pos[:, -1:] + 1

    # -- public API --------------------------------------------------------------------------------
    @torch.inference_mode()
    def generate_stream(self, text: str, speaker: int, context: List[Segment], max_audio_length_ms: float = 90_000,
                        temperature: float = 0.7, topk: int = 30,
                        on_chunk_generated: Optional[Callable[[torch.Tensor], None]] = None) -> Iterator[torch.Tensor]:
        buf: List[torch.Tensor] = []
        stream = None
        if self._stream_stateful and hasattr(self._audio_tokenizer, "streaming"):
            stream = self._audio_tokenizer.streaming()

        def decode(frames: List[torch.Tensor]) -> torch.Tensor:
            if stream is None:
                return self._decode_frames(frames)
            return stream.decode(torch.stack(frames).permute(1, 2, 0)).squeeze(0).squeeze(0)

        for sample in self._frames(text, speaker, context, max_audio_length_ms, temperature, topk):
            buf.append(sample)
            if len(buf) >= self._stream_buffer_size:
                chunk = decode(buf)
                buf = []
                if on_chunk_generated:
                    on_chunk_generated(chunk)
                yield chunk
        if buf:
            chunk = decode(buf)
            if on_chunk_generated:
                on_chunk_generated(chunk)
            yield chunk

    @torch.inference_mode()
    def generate(self, text: str, speaker: int, context: List[Segment], max_audio_length_ms: float = 90_000,
                 temperature: float = 0.7, topk: int = 30, stream: bool = False) -> torch.Tensor:
        if stream:
            chunks = list(self.generate_stream(text, speaker, context, max_audio_length_ms, temperature, topk))
            return torch.cat(chunks) if chunks else torch.tensor([])
        samples = list(self._frames(text, speaker, context, max_audio_length_ms, temperature, topk))
        return self._decode_frames(samples)


class AudioStreamWriter:
    """Collects streamed chunks and writes them to one file (reference ``generator.py:303-327``)."""

    def __init__(self, filename, sample_rate):
        self.filename = filename
        self.sample_rate = sample_rate
        self.audio_chunks: List[torch.Tensor] = []
        self.lock = threading.Lock()

    def add_chunk(self, chunk):
        with self.lock:
            self.audio_chunks.append(chunk)

    def write_file(self):
        with self.lock:
            if not self.audio_chunks:
                return
            audio = torch.cat(self.audio_chunks)
            _save_wav(self.filename, audio.unsqueeze(0).cpu(), self.sample_rate)


def _save_wav(filename, audio: torch.Tensor, sample_rate: int) -> None:
    """``torchaudio.save`` as the reference calls it; without torchaudio, a 16-bit PCM file via ``wave``."""
    try:
        import torchaudio

        torchaudio.save(filename, audio, sample_rate)
        return
    except ImportError:
        pass
    import wave

    pcm = (audio.float().clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16)
    with wave.open(str(filename), "wb") as f:
        f.setnchannels(pcm.shape[0])
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.t().contiguous().numpy().tobytes())


def load_csm_1b(device: str = "cuda", *, model_path: str = "sesame/csm-1b", text_tokenizer=None,
                audio_tokenizer: Optional[MimiCodec] = None) -> Generator:
    """Reference ``load_csm_1b(device)`` (``generator.py:330-346``) minus the cuDNN / torch.compile knobs, which
    have no counterpart here: weights through ``Model.from_pretrained`` (the hub repo, or -- keyword extras for
    offline use -- a local directory written by ``save_pretrained``), bf16 on ``device``, caches for batch 1."""
    model = Model.from_pretrained(model_path)
    model.to(device=device, dtype=torch.bfloat16)
    return Generator(model, text_tokenizer=text_tokenizer, audio_tokenizer=audio_tokenizer)


def generate_streaming_audio(generator: Generator, text: str, speaker: int, context: List[Segment], output_file: str,
                             max_audio_length_ms: float = 90_000, temperature: float = 0.7, topk: int = 30,
                             play_audio: bool = False):
    """Reference ``generate_streaming_audio`` (``generator.py:349-433``): stream chunks into a file, optionally
    playing them as they arrive (needs ``sounddevice``)."""
    writer = AudioStreamWriter(output_file, generator.sample_rate)
    audio_queue: "queue.Queue[torch.Tensor]" = queue.Queue()
    stop_event = threading.Event()
    player_thread = None
    if play_audio:
        try:
            import sounddevice as sd

            def audio_player():
                while not stop_event.is_set() or not audio_queue.empty():
                    try:
                        chunk = audio_queue.get(timeout=0.5)
                    except queue.Empty:
                        continue
                    sd.play(chunk.cpu().numpy(), generator.sample_rate)
                    sd.wait()

            player_thread = threading.Thread(target=audio_player)
            player_thread.start()
        except ImportError:
            print("sounddevice library not found. Install with 'pip install sounddevice' to enable real-time playback.")
            play_audio = False

    def on_chunk_generated(chunk):
        writer.add_chunk(chunk)
        if play_audio:
            audio_queue.put(chunk)

    print("Generating audio in streaming mode...")
    start_time = time.time()
    chunk_count = 0
    for _ in generator.generate_stream(text=text, speaker=speaker, context=context, max_audio_length_ms=max_audio_length_ms,
                                       temperature=temperature, topk=topk, on_chunk_generated=on_chunk_generated):
        chunk_count += 1
        print(f"Generated chunk {chunk_count}")
    writer.write_file()
    if player_thread is not None:
        stop_event.set()
        player_thread.join()
    print(f"Audio generation completed in {time.time() - start_time:.2f} seconds")
