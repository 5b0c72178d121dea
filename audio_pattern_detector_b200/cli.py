"""`audio-pattern-detector` command line (same sub-commands and flags as reference cli.py:19-67)."""
import argparse
import sys


def _match(args: argparse.Namespace) -> None:
    from .match import cmd_match          # lazy: keeps `--help` fast
    cmd_match(args)


def _show_config(args: argparse.Namespace) -> None:
    from .match import cmd_show_config
    cmd_show_config(args)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="audio-pattern-detector", description="Audio pattern detection tools")
    sub = parser.add_subparsers(dest="command", help="Available commands")

    m = sub.add_parser("match", help="Find pattern matches in audio files")
    m.add_argument("--pattern-file", metavar="pattern file", type=str, action="append",
                   help="pattern file (can be specified multiple times)")
    m.add_argument("--pattern-folder", metavar="pattern folder", type=str, action="append",
                   help="folder with pattern audio clips (can be specified multiple times, "
                        "can be combined with --pattern-file)")
    m.add_argument("audio_file", nargs="?", default=None, type=str,
                   help="single audio file to find pattern in (omit when using --stdin or --multiplexed-stdin)")
    m.add_argument("--stdin", action="store_true", help="read audio from stdin in WAV format")
    m.add_argument("--multiplexed-stdin", action="store_true",
                   help="read patterns and audio from stdin using multiplexed protocol (always outputs JSONL). "
                        "Protocol: [uint32 num_patterns] then for each pattern "
                        "[uint32 name_len][name][uint32 data_len][wav_data], followed by audio stream (WAV)")
    m.add_argument("--target-sample-rate", metavar="rate", type=int,
                   help="target sample rate for processing in Hz (default: 8000)")
    m.add_argument("--timestamp-format", choices=["ms", "formatted", "both"], default="both",
                   help='timestamp format in JSONL output: "both" (default), "ms" or "formatted"')
    m.add_argument("--chunk-seconds", metavar="seconds", type=str, default="60",
                   help='seconds per chunk for sliding window (default: 60, "auto" derives it from the patterns)')
    m.add_argument("--debug", action=argparse.BooleanOptionalAction, default=False,
                   help="debug mode (audio file only)")
    m.add_argument("--debug-dir", metavar="dir", type=str, default="./tmp",
                   help="base directory for debug output (default: ./tmp)")
    m.add_argument("--height-min", metavar="height", type=float, default=None,
                   help="override minimum correlation peak height (default: 0.25)")
    m.set_defaults(func=_match)

    s = sub.add_parser("show-config", help="Show computed configuration for a pattern file")
    s.add_argument("pattern_file", metavar="pattern file", type=str, help="pattern file")
    s.add_argument("--target-sample-rate", metavar="rate", type=int,
                   help="target sample rate for processing in Hz (default: 8000)")
    s.set_defaults(func=_show_config)
    return parser


def main() -> None:
    parser = build_parser()
    args = parser.parse_args()
    if not args.command:
        parser.print_help()
        sys.exit(1)
    args.func(args)


if __name__ == "__main__":
    main()
