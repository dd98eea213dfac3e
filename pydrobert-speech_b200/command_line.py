"""Command-line front ends: ``signals-to-torch-feat-dir`` and ``compute-feats-from-kaldi-tables``.

Flags, file formats, return codes and the seeding contract are the reference's
(``pydrobert/speech/command_line.py``); the processing loop is not.  The reference pushes one
utterance at a time through a ``DataLoader`` of CPU workers.  Here reader threads decode the
signals, utterances are packed into batches of ``--batch-samples`` samples and each batch goes
through :class:`..pipeline.FeaturePipeline` (one fused launch per batch, copies overlapped with
compute); writer threads store the per-utterance ``.pt`` files and append to the manifest only
after a file is on disk, so an interrupted run resumes exactly like the reference's.

Several GPUs: launched under ``torchrun`` (``python -m torch.distributed.run --nproc-per-node N -m
pydrobert_speech_b200.command_line signals-to-torch-feat-dir ...``) every rank takes a shard of
the utterances (``pipeline.shard_utterances`` on the file sizes), computes it on its own GPU and
writes its own files; there is no data-path communication.  The reference's counterpart is
``DataLoader(num_workers=N)`` (``command_line.py:585-594``).

``--seed`` determinism: the dither stream is keyed by ``(seed, utterance index in the map,
sample)``, the batched analogue of the reference's ``torch.manual_seed(seed + idx)``; results do
not depend on ``--num-workers``, on the batch size or on the number of ranks.
"""

import argparse
import collections
import json
import logging
import os
import sys
import threading
import time

from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np

from . import config
from .alias import alias_factory_subclass_from_arg
from .compute import FrameComputer, SIFrameComputer, STFTFrameComputer
from .post import PostProcessor
from .pre import Dither, Preemphasize, PreProcessor
from .util import read_signal

__all__ = ["compute_feats_from_kaldi_tables", "signals_to_torch_feat_dir", "main"]

try:
    from ruamel.yaml import YAML

    def _load_config(string: str):
        return YAML(typ="safe").load(string)

    _HAVE_YAML = True
except ImportError:
    from json import loads as _load_config

    _HAVE_YAML = False

_EPILOGUE = """JSON arguments may be given inline or as a path to a file.  If ruamel.yaml is installed they
are parsed as YAML 1.2 (of which JSON is a subset)."""


def _config_type(string: str):
    """A JSON/YAML string, or a path to a file holding one, as a container hierarchy"""
    name = string
    try:
        with open(string) as handle:
            string = handle.read()
    except IOError:
        pass
    try:
        return _load_config(string)
    except Exception as e:
        msg = f"Unable to parse '{name}' as JSON" + (" or YAML" if _HAVE_YAML else "")
        if not _HAVE_YAML and name.endswith(".yaml"):
            msg += ". This could be a YAML file. Install ruamel.yaml to try it"
        raise ValueError(msg) from e


def _nonneg_int_type(string):
    try:
        val = int(string)
        assert val >= 0
    except (ValueError, AssertionError):
        raise argparse.ArgumentTypeError("{} is not a nonnegative integer".format(string))
    return val


def _build_list(base, spec):
    """``--preprocess`` / ``--postprocess`` accept one config or a list of them"""
    specs = [spec] if isinstance(spec, dict) else list(spec)
    return [alias_factory_subclass_from_arg(base, element) for element in specs]


def _select_channel(signal: np.ndarray, channel: int, utt_id: str) -> np.ndarray:
    # same checks and messages as command_line.py:112-125 of the reference
    if channel == -1 and signal.ndim > 1 and signal.shape[0] > 1:
        raise ValueError(
            "Utterance {}: Channel is not specified but signal has shape {}".format(utt_id, signal.shape)
        )
    if (channel != -1 and signal.ndim == 1) or (channel >= signal.shape[0]):
        raise ValueError(
            "Utterance {}: Channel specified as {} but signal has shape {}".format(
                utt_id, channel, signal.shape
            )
        )
    return signal if signal.ndim == 1 else signal[channel]


# ----------------------------------------------------------------------------------------------
# signals-to-torch-feat-dir
# ----------------------------------------------------------------------------------------------
def _signals_to_torch_feat_dir_parse_args(args):
    parser = argparse.ArgumentParser(
        prog="signals-to-torch-feat-dir",
        description=signals_to_torch_feat_dir.__doc__,
        formatter_class=argparse.RawDescriptionHelpFormatter,
        epilog=_EPILOGUE,
    )
    parser.add_argument("map", type=argparse.FileType("r"),
                        help="Path to the file containing (<utterance>, <path>) pairs")
    parser.add_argument("computer_config", type=_config_type, nargs="?", default=None,
                        help="JSON file or string configuring the FrameComputer. If unspecified, the "
                        "audio (with channels removed) is stored directly with shape (S, 1)")
    parser.add_argument("dir", help="Directory to output features to (created if missing)")
    parser.add_argument("--channel", type=int, default=-1,
                        help="Channel to draw audio from. Default is to assume mono")
    parser.add_argument("--preprocess", type=_config_type, default=tuple(),
                        help="JSON list of PreProcessor configurations, applied in order")
    parser.add_argument("--postprocess", type=_config_type, default=tuple(),
                        help="JSON list of PostProcessor configurations, applied in order")
    parser.add_argument(
        "--force-as", default=None,
        choices={"table", "wav", "hdf5", "npy", "npz", "pt", "sph", "kaldi", "file", "soundfile"}
        | config.SOUNDFILE_SUPPORTED_FILE_TYPES,
        help="Force the paths in 'map' to be interpreted as a specific type of data")
    parser.add_argument("--seed", type=_nonneg_int_type, default=None,
                        help="Seed for operations like dithering. If unset, one is drawn at random")
    parser.add_argument("--file-prefix", default="", help="Prefix of the output file names")
    parser.add_argument("--file-suffix", default=".pt", help="Suffix of the output file names")
    parser.add_argument("--num-workers", type=_nonneg_int_type, default=0,
                        help="Threads decoding signal files. Never affects the results")
    parser.add_argument("--manifest", type=argparse.FileType("a+"), default=None,
                        help="File listing the utterances already computed; they are skipped and new "
                        "ones appended, so that an interrupted run can be resumed")
    parser.add_argument("--batch-samples", type=_nonneg_int_type, default=1 << 24,
                        help="Samples per GPU batch (an implementation knob of this build)")
    parser.add_argument("--report", default=None,
                        help="Append one JSON line of throughput figures (files/s, audio-hours/s including "
                        "decoding and torch.save) to this file (an addition of this build)")
    return parser.parse_args(args)


def signals_to_torch_feat_dir(args: Optional[Sequence[str]] = None) -> int:
    """Convert a map of signals to a torch SpectDataSet

    Reads a text file of "<utt_id> <path_to_signal>" lines, computes features according to the
    passed-in settings, and stores each utterance's features as a torch.FloatTensor of shape
    (T, F) in "dir/<file_prefix><utt_id><file_suffix>".

    Signals are read with "util.read_signal()" and are expected to have shape (C, S), or (S,) when
    "--channel" is -1.  No check is made that the signals match the computer's sampling rate.
    """
    try:
        options = _signals_to_torch_feat_dir_parse_args(args)
    except SystemExit as ex:
        return ex.code
    try:
        import torch
    except ImportError:
        print("signals-to-torch-feat-dir requires a PyTorch installation", file=sys.stderr)
        return 1
    from .pipeline import FeaturePipeline

    seed = np.random.randint(np.iinfo(np.int32).max) if options.seed is None else options.seed
    utt2path = dict()
    for line_no, line in enumerate(options.map):
        line = line.strip()
        if not line:
            continue
        fields = line.split(" ")
        if len(fields) < 2:
            print("Line {} of {}: not of format <utt_id> <path>".format(line_no + 1, options.map.name),
                  file=sys.stderr)
            return 1
        if fields[0] in utt2path:
            print('Line {} of {}: "{}" already exists as utterance'.format(
                line_no + 1, options.map.name, fields[0]), file=sys.stderr)
            return 1
        utt2path[fields[0]] = " ".join(fields[1:])
    # the dither stream is keyed by the position in the *full* map, before the manifest filter and
    # before sharding: the features do not depend on --num-workers, on the batch size, on which
    # utterances were already done, or on the number of GPUs
    utt_index = {utt: idx for idx, utt in enumerate(utt2path)}
    done = set()
    if options.manifest is not None:
        options.manifest.seek(0)
        done = {line.strip() for line in options.manifest}

    # ---- one process per GPU under torchrun: every rank takes a shard of the utterances ----------
    rank, world, local_rank = _dist_env()
    if world > 1 and options.seed is None and options.preprocess:
        print("--seed must be given when pre-processing on more than one rank (every rank would "
              "draw a seed of its own)", file=sys.stderr)
        return 1
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % max(1, torch.cuda.device_count()))
    items = list(utt2path.items())
    if world > 1:
        # the shards are cut from the FULL map, before the manifest filter: every rank derives the
        # same partition no matter what the other ranks have already appended to the manifest
        from .pipeline import shard_utterances

        sizes = [_path_size(path) for _, path in items]
        items = [items[i] for i in shard_utterances(sizes, world)[rank]]
    items = [item for item in items if item[0] not in done]

    computer = None
    if options.computer_config is not None:
        computer = alias_factory_subclass_from_arg(FrameComputer, options.computer_config)
        if not isinstance(computer, (STFTFrameComputer, SIFrameComputer)):
            raise NotImplementedError
    preprocessors = _build_list(PreProcessor, options.preprocess)
    if not all(isinstance(p, (Dither, Preemphasize)) for p in preprocessors):
        raise NotImplementedError
    postprocessors = _build_list(PostProcessor, options.postprocess)
    pipeline = None
    if computer is not None:
        # post_along_time=False: the reference applies post-processors with their default axis
        pipeline = FeaturePipeline(computer, preprocessors, postprocessors, seed=seed,
                                   chunk_samples=options.batch_samples, post_along_time=False)

    def load(item):
        utt_id, path = item
        try:
            # native sample type: 16-bit PCM goes to the GPU as it is (the reference's float64 copy of
            # it holds the same values); everything else is computed in float32 anyway
            signal = read_signal(path, dtype=None, force_as=options.force_as, key=utt_id)
            if signal.dtype != np.int16:
                signal = signal.astype(np.float32 if pipeline is not None else np.float64, copy=False)
            elif pipeline is None:
                signal = signal.astype(np.float64)
        except Exception as e:
            raise IOError(f"Utterance {utt_id}: {e}") from e
        return _select_channel(signal, options.channel, utt_id)

    os.makedirs(options.dir, exist_ok=True)
    manifest_lock = threading.Lock()
    stats = dict(utterances=0, samples=0, frames=0)

    def write(utt_id, feats):
        torch.save(torch.as_tensor(np.ascontiguousarray(feats)).float(),
                   os.path.join(options.dir, options.file_prefix + utt_id + options.file_suffix))
        if options.manifest is not None:
            with manifest_lock:  # one short line per write(2) on an O_APPEND file: safe across ranks too
                options.manifest.write(utt_id + "\n")
                options.manifest.flush()

    def compute(batch):
        utts, signals = zip(*batch)
        if pipeline is None:
            # raw passthrough (S, 1); pre-processors seeded per utterance like the reference
            from .torch import PyTorchDither, PyTorchPreemphasize

            mods = [PyTorchDither.from_dither(p) if isinstance(p, Dither)
                    else PyTorchPreemphasize.from_preemphasize(p) for p in preprocessors]
            outs = []
            for utt_id, signal in zip(utts, signals):
                torch.manual_seed(seed + utt_index[utt_id])
                tensor = torch.from_numpy(np.ascontiguousarray(signal))
                for mod in mods:
                    tensor = mod(tensor)
                outs.append(tensor.unsqueeze(1).numpy())
            return utts, outs
        # every utterance keeps the dither stream of its position in the map
        outs = []
        run_start = 0
        for i in range(1, len(utts) + 1):
            if i == len(utts) or utt_index[utts[i]] != utt_index[utts[i - 1]] + 1:
                outs.extend(pipeline.run_list(signals[run_start:i], utt_base=utt_index[utts[run_start]]))
                run_start = i
        return utts, outs

    # Decoder threads run at most two batches ahead of the GPU (a sliding window of futures: the
    # decoded signals of a large corpus must not pile up in host memory), torch.save runs on writer
    # threads so that the GPU batch of the next utterances never waits for the file system.
    workers = max(1, options.num_workers)
    budget = 2 * max(1, options.batch_samples)
    t_begin = time.perf_counter()
    pending_writes = collections.deque()
    with ThreadPoolExecutor(workers) as readers, ThreadPoolExecutor(max(2, workers)) as writers:
        window = collections.deque()  # (item, future, size estimate)
        in_flight = 0
        position = 0
        batch, batch_samples = [], 0

        def drain_writes(limit):
            while len(pending_writes) > limit:
                pending_writes.popleft().result()

        def flush():
            nonlocal batch, batch_samples
            if batch:
                utts, outs = compute(batch)
                for utt_id, feats in zip(utts, outs):
                    stats["frames"] += len(feats)
                    pending_writes.append(writers.submit(write, utt_id, feats))
                drain_writes(4096)
            batch, batch_samples = [], 0

        while position < len(items) or window:
            while position < len(items) and (in_flight < budget or len(window) < workers):
                item = items[position]
                estimate = max(1, _path_size(item[1]) // 2)
                window.append((item, readers.submit(load, item), estimate))
                in_flight += estimate
                position += 1
            item, future, estimate = window.popleft()
            signal = future.result()
            in_flight -= estimate
            batch.append((item[0], signal))
            batch_samples += len(signal)
            stats["utterances"] += 1
            stats["samples"] += len(signal)
            if batch_samples >= max(1, options.batch_samples):
                flush()
        flush()
        drain_writes(0)
    if options.report is not None:
        elapsed = time.perf_counter() - t_begin
        rate = float(computer.sampling_rate) if computer is not None else 16000.0
        line = dict(rank=rank, world_size=world, seconds=elapsed, utterances=stats["utterances"],
                    files_per_second=stats["utterances"] / elapsed if elapsed else None,
                    audio_hours=stats["samples"] / rate / 3600.0,
                    audio_hours_per_second=stats["samples"] / rate / 3600.0 / elapsed if elapsed else None,
                    frames=stats["frames"])
        with open(options.report, "a") as handle:
            handle.write(json.dumps(line) + "\n")
    return 0


def _dist_env():
    """(rank, world size, local rank) of a torchrun / torch.distributed.run launch, else (0, 1, 0)"""
    try:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    except ValueError:
        return 0, 1, 0
    if world < 1 or not 0 <= rank < world:
        return 0, 1, 0
    return rank, world, local_rank


def _path_size(path: str) -> int:
    try:
        return os.path.getsize(path)
    except OSError:
        return 1


# ----------------------------------------------------------------------------------------------
# compute-feats-from-kaldi-tables
# ----------------------------------------------------------------------------------------------
def _kaldi_parse_args(args, logger):
    """The reference builds its parser from pydrobert-kaldi's ``KaldiParser`` (``command_line.py:179-240``):
    plain argparse plus ``-v/--verbose`` and ``--config <file>`` (one ``--option=value`` per line)"""
    parser = argparse.ArgumentParser(
        prog="compute-feats-from-kaldi-tables", description=compute_feats_from_kaldi_tables.__doc__,
        formatter_class=argparse.RawDescriptionHelpFormatter, epilog=_EPILOGUE)
    parser.add_argument("wav_rspecifier", help="Input wave table rspecifier")
    parser.add_argument("feats_wspecifier", help="Output feature table wspecifier")
    parser.add_argument("computer_config", type=_config_type,
                        help="JSON file or string configuring the FrameComputer")
    parser.add_argument("-v", "--verbose", type=int, default=0, help="Verbose level (higher->more logging)")
    parser.add_argument("--config", default=None, help="File of additional '--option=value' lines")
    parser.add_argument("--min-duration", type=float, default=0.0,
                        help="Min duration of segments to process (in seconds)")
    parser.add_argument("--channel", type=int, default=-1,
                        help="Channel to draw audio from. Default is to assume mono")
    parser.add_argument("--preprocess", type=_config_type, default=tuple())
    parser.add_argument("--postprocess", type=_config_type, default=tuple())
    parser.add_argument("--seed", type=_nonneg_int_type, default=None)
    parser.add_argument("--batch-samples", type=_nonneg_int_type, default=1 << 24,
                        help="Samples per GPU batch (an implementation knob of this build)")
    args = list(sys.argv[1:] if args is None else args)
    for i, arg in enumerate(args):  # options of a Kaldi-style config file come first; the command line wins
        if arg.startswith("--config="):
            path = arg.split("=", 1)[1]
        elif arg == "--config" and i + 1 < len(args):
            path = args[i + 1]
        else:
            continue
        with open(path) as handle:
            extra = [line.split("#", 1)[0].strip() for line in handle]
        args = [line for line in extra if line] + args
        break
    options = parser.parse_args(args)
    logger.setLevel(logging.DEBUG if options.verbose > 0 else (logging.INFO if options.verbose == 0 else logging.WARNING))
    return options


def compute_feats_from_kaldi_tables(args: Optional[Sequence[str]] = None) -> int:
    """Store features from a kaldi archive in a kaldi archive

    This command is intended to replace Kaldi's (https://kaldi-asr.org/) series of
    "compute-<something>-feats" scripts in a Kaldi pipeline.

    Drop-in for the reference command of the same name (``command_line.py:245-359``): same
    arguments, warnings and return codes.  Wave and feature tables are read and written by
    pydrobert-kaldi when it is installed and by the built-in ``_kaldi_io`` otherwise (``ark:`` /
    ``scp:`` files, binary or text matrices, no pipes).  Utterances are computed in batches of
    ``--batch-samples`` samples; as in the reference the post-processors are validated but not
    applied.
    """
    logger = logging.getLogger(sys.argv[0])
    if not logger.handlers:
        logger.addHandler(logging.StreamHandler())
    try:
        options = _kaldi_parse_args(args, logger)
    except SystemExit as ex:
        return ex.code
    from .pipeline import FeaturePipeline

    try:
        computer = alias_factory_subclass_from_arg(FrameComputer, options.computer_config)
    except ValueError:
        logger.error("Failed to build computer:", exc_info=True)
        return 1
    try:
        preprocessors = _build_list(PreProcessor, options.preprocess)
    except ValueError:
        logger.error("Failed to build preprocessor:", exc_info=True)
        return 1
    try:
        _build_list(PostProcessor, options.postprocess)  # validated but, as in the reference, unused
    except ValueError:
        logger.error("Failed to build postprocessor:", exc_info=True)
        return 1
    seed = np.random.randint(np.iinfo(np.int32).max) if options.seed is None else options.seed
    pipeline = FeaturePipeline(computer, preprocessors, seed=seed, chunk_samples=options.batch_samples)
    try:
        from pydrobert.kaldi.io import open as kaldi_open  # type: ignore
        from pydrobert.kaldi.io.enums import KaldiDataType  # type: ignore

        open_waves = lambda spec: kaldi_open(spec, "wm", value_style="bsd")  # noqa: E731
        open_feats = lambda spec: kaldi_open(spec, "bm", mode="w")  # noqa: E731
        as_double = bool(KaldiDataType.BaseMatrix.is_double)
    except ImportError:
        from ._kaldi_io import MatrixTableWriter, WaveTableReader

        open_waves, open_feats, as_double = WaveTableReader, MatrixTableWriter, False
    try:
        wav_reader = open_waves(options.wav_rspecifier)
    except IOError:
        logger.error("Could not read the wave table {}".format(options.wav_rspecifier))
        return 1
    try:
        feat_writer = open_feats(options.feats_wspecifier)
    except IOError:
        logger.error("Could not open the feat table {} for writing".format(options.feats_wspecifier))
        return 1
    num_utts, num_success = 0, 0
    batch, batch_samples = [], 0  # (utt_id, position in the table, signal)

    def flush():
        nonlocal batch, batch_samples, num_success
        run_start = 0
        for i in range(1, len(batch) + 1):  # runs of consecutive positions keep their dither streams
            if i == len(batch) or batch[i][1] != batch[i - 1][1] + 1:
                feats = pipeline.run_list([sig for _, _, sig in batch[run_start:i]], utt_base=batch[run_start][1])
                for (utt_id, _, _), feat in zip(batch[run_start:i], feats):
                    feat_writer.write(utt_id, feat.astype(np.float64) if as_double else feat)
                    num_success += 1
                run_start = i
        batch, batch_samples = [], 0

    for utt_id, (buff, samp_freq, duration) in wav_reader.items():
        num_utts += 1
        if duration < options.min_duration:
            logger.warning("File: {} is too short ({:.2f} sec): producing no output".format(utt_id, duration))
            continue
        if samp_freq != computer.sampling_rate:
            logger.warning("Sample frequency mismatch for file {}: you specified {:.2f} but data has "
                           "{:.2f}: producing no output".format(utt_id, computer.sampling_rate, samp_freq))
            continue
        channel = options.channel
        if channel == -1 and buff.shape[0] > 1:
            logger.warning("Channel is not specified but you have data with {} channels; defaulting "
                           "to zero".format(buff.shape[0]))
            channel = 0
        elif channel >= buff.shape[0]:
            logger.warning("File with id {} has {} channels but you specified channel {}, producing no "
                           "output".format(utt_id, buff.shape[0], channel))
            continue
        signal = np.ascontiguousarray(buff[max(channel, 0)], dtype=np.float32)
        batch.append((utt_id, num_utts - 1, signal))
        batch_samples += len(signal)
        if batch_samples >= max(1, options.batch_samples):
            flush()
        if num_utts % 10 == 0:
            logger.info("Processed {} utterances".format(num_utts))
    flush()
    logger.info("Done {} out of {} utterances".format(num_success, num_utts))
    feat_writer.close()
    wav_reader.close()
    return 0 if num_success else 1


def main(argv: Optional[Sequence[str]] = None) -> int:
    """``python -m pydrobert_speech_b200.command_line <command> ...``"""
    argv = list(sys.argv[1:] if argv is None else argv)
    commands = {
        "signals-to-torch-feat-dir": signals_to_torch_feat_dir,
        "compute-feats-from-kaldi-tables": compute_feats_from_kaldi_tables,
    }
    if not argv or argv[0] not in commands:
        print("usage: python -m pydrobert_speech_b200.command_line {%s} ..." % ",".join(commands),
              file=sys.stderr)
        return 2
    return commands[argv[0]](argv[1:])


if __name__ == "__main__":
    sys.exit(main())
