// filterbank.cu — placeholder
