"""pytest glue of the reference's own test-suite: not needed to run the step loop."""
