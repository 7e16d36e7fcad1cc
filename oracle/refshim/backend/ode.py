"""Reference solutions / example problems of the reference's test-suite: not needed to run the step loop."""
