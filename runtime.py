"""GPU runtime check (same command line as the reference's runtime.py): per validation image, the time of
`model.fwd_runtime` between device synchronisations.  Implementation: larvanet_b200/entrypoints.py."""
from larvanet_b200.entrypoints import runtime_main as main

if __name__ == '__main__':
    main()
