"""Command-line entry points (``unicore-train`` -> :func:`unicore_cli.train.cli_main`)."""
