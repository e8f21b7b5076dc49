"""Reference arm of the benchmark (NOT product code): locates and drives the UNMODIFIED reference implementation
(jerryyang1231/whisper-flamingo) for timing.  See baseline/reference_arm.py."""
