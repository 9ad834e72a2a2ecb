"""Device timing, clock sampling and build helpers."""
