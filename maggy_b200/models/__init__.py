"""Model zoo on the maggy_b200 kernels: GPT-2 (flagship), Llama-3, BERT (MLM), ResNet-50."""


def __getattr__(name):
    import importlib

    table = {"GPT2": "gpt2", "GPT2Config": "gpt2", "TorchGPT2": "gpt2", "Llama": "llama", "LlamaConfig": "llama",
             "BertForMaskedLM": "bert", "BertConfig": "bert", "ResNet": "resnet", "resnet50": "resnet"}
    if name in table:
        return getattr(importlib.import_module(f"{__name__}.{table[name]}"), name)
    raise AttributeError(name)
