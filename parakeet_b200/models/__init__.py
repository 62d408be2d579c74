from .parallel_wavegan import PWGGenerator, PWGInference  # noqa: F401
