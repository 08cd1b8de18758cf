"""Drop-in alias package: lets ``python -m pht.train -cn <cfg>`` and
``from pht.models.afgsa.model import AFGSANet`` resolve to the B200 implementation."""
